! dmdb200_iso_c.f90 -- Fortran 2003 ISO_C_BINDING interface to libdmdb200.so (include/dmdb200.h).
!
! This is the host-language side BASELINE.json's north_star asks for: the reference program (code/main.F90
! of HallandSantiso-NCSU/Parallel-DMD-for-biomolecules) keeps its Fortran set-up and file I/O and calls the
! B200 engine through these bind(C) interfaces instead of its own events()/nbor()/main loop.  Every interface
! names the reference routine it replaces (paths relative to parallel-dmd-PRIME20/code).
!
! STATUS: kept as source.  Neither this container nor the GPU boxes have a Fortran compiler (gfortran, ifort,
! nvfortran, flang: not found -- SURVEY.md 8c), so this file is NOT compiled by __graft_entry__.build(); the
! same symbols are exercised through the C ABI by the ctypes binding (parallel_dmd_for_biomolecules_b200/dmd.py)
! and tests/test_capi_symbols.py checks that every name bound here is exported by the shared library.
! Build where a compiler exists:   gfortran -c dmdb200_iso_c.f90 && gfortran dmd_b200_main.f90 dmdb200_iso_c.o \
!                                   -L.. -ldmdb200 -Wl,-rpath,'$ORIGIN/..' -o dmd_b200
module dmdb200
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: dmdb_params, dmdb_tables, dmdb_topology, dmdb_stats, dmdb_energy, dmdb_event
  public :: dmdb_create, dmdb_destroy, dmdb_last_error, dmdb_num_beads, dmdb_num_cells
  public :: dmdb_set_state, dmdb_set_state_all, dmdb_get_state_all, dmdb_set_temperature
  public :: dmdb_nbor, dmdb_predict_all, dmdb_run, dmdb_run_until_output, dmdb_sync_positions
  public :: dmdb_get_cells, dmdb_get_nbors, dmdb_get_calendar, dmdb_get_state, dmdb_get_evcode
  public :: dmdb_energy_of, dmdb_get_event_log, dmdb_get_replica_stats
  public :: dmdb_potential_energies, dmdb_apply_temperatures, dmdb_get_batch_stats
  public :: dmdb_device_fill, dmdb_set_service_ctas
  public :: dmdb_exchange_stats, dmdb_nccl_unique_id, dmdb_comm_init, dmdb_exchange, dmdb_exchange_gathered
  public :: dmdb_sheet_observables
  public :: dmdb_error_message
  public :: DMDB_OK, DMDB_ERR_ARG, DMDB_ERR_NO_DEVICE, DMDB_ERR_CUDA, DMDB_ERR_STATE, DMDB_ERR_CAPACITY, &
            DMDB_ERR_PHYSICS, DMDB_MAX_SPECIES

  integer(c_int), parameter :: DMDB_OK = 0, DMDB_ERR_ARG = 1, DMDB_ERR_NO_DEVICE = 2, DMDB_ERR_CUDA = 3, &
                               DMDB_ERR_STATE = 4, DMDB_ERR_CAPACITY = 5, DMDB_ERR_PHYSICS = 6
  integer, parameter :: DMDB_MAX_SPECIES = 2

  ! raw parameter-file contents exactly as inputinfo.f:162-404 reads them
  type, bind(C) :: dmdb_tables
    real(c_double) :: protein(12)     ! parameters/protein.data
    real(c_double) :: ep(400)         ! parametersep/ep19p_ha55a_weakhp.data, (i-9)*20+(j-9)+1, file sign
    real(c_double) :: bds(400)        ! parameters/beadwell_ha55a.data bead diameters
    real(c_double) :: wel(400)        ! parameters/beadwell_ha55a.data well diameters
    real(c_double) :: mass(28)        ! parameters/mass.data by identity id
    real(c_double) :: rcarnrco(120)   ! parameters/rcarnrco.data, 20 rows x 6 (row-major)
    real(c_double) :: sqz6to10(100)   ! parameters/sqz6to10.data, 20 rows x 5 in file column order
  end type

  ! what the reference fixes with -Dnop1 -Dnop2 -Dchnln1 -Dchnln2 -Dnumbeads1 -Dnumbeads2 (qfile/script.sh:7)
  type, bind(C) :: dmdb_topology
    integer(c_int32_t) :: n_species
    integer(c_int32_t) :: n_chains(DMDB_MAX_SPECIES)
    integer(c_int32_t) :: chnln(DMDB_MAX_SPECIES)
    integer(c_int32_t) :: numbeads(DMDB_MAX_SPECIES)
    type(c_ptr) :: identity(DMDB_MAX_SPECIES)   ! c_loc of identity.inp rows of each species
    type(c_ptr) :: hp(DMDB_MAX_SPECIES)         ! c_loc of hp1.inp / hp2.inp
    type(c_ptr) :: firstside(DMDB_MAX_SPECIES)  ! c_loc of firstside1.data / firstside2.data
  end type

  ! the two stdin numbers (main.F90:126-128), the box length (inputinfo.f:78) and the -D behaviour flags
  type, bind(C) :: dmdb_params
    real(c_double) :: boxl, tstar
    integer(c_int32_t) :: canon, no_hbs, n_wrap, n_replicas, device, nbr_capacity, log_capacity, engine
    integer(c_int64_t) :: seed   ! uint64_t in C; same bits
  end type

  type, bind(C) :: dmdb_event
    real(c_double) :: t
    integer(c_int32_t) :: i, j, type, evcode
  end type

  type, bind(C) :: dmdb_stats     ! main.F90:1356-1363 tallies
    integer(c_int64_t) :: events, pair_events
    integer(c_int64_t) :: nevents(32)
    integer(c_int64_t) :: ghosts, updates, forced_updates, pair_predictions, nbr_visits
    real(c_double) :: device_ms
    integer(c_int32_t) :: kernel_launches, reserved
  end type

  type, bind(C) :: dmdb_exchange_stats   ! one replica-exchange step (dmdb_exchange)
    integer(c_int32_t) :: ladders, attempted, accepted, changed_local
    real(c_double) :: device_ms
    integer(c_int32_t) :: kernel_launches, reserved
  end type

  type, bind(C) :: dmdb_energy    ! energy.f:25-101 outputs
    real(c_double) :: ered, tred, sumvel, ehh_ii, ehh_ij
    integer(c_int32_t) :: hb_alpha, hb_ii, hb_ij, reserved
  end type

  interface
    ! program start-up main.F90:117-234 (+ inputinfo/scale_down/make_code/nbor_setup done inside the library)
    function dmdb_create(p, topo, tab, handle) bind(C, name="dmdb_create") result(rc)
      import :: c_int, c_ptr, dmdb_params, dmdb_topology, dmdb_tables
      type(dmdb_params), intent(in) :: p
      type(dmdb_topology), intent(in) :: topo
      type(dmdb_tables), intent(in) :: tab
      type(c_ptr), intent(out) :: handle
      integer(c_int) :: rc
    end function
    subroutine dmdb_destroy(handle) bind(C, name="dmdb_destroy")
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine
    function dmdb_last_error(handle) bind(C, name="dmdb_last_error") result(msg)
      import :: c_ptr
      type(c_ptr), value :: handle
      type(c_ptr) :: msg
    end function
    function dmdb_num_beads(handle) bind(C, name="dmdb_num_beads") result(n)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: n
    end function
    function dmdb_num_cells(handle) bind(C, name="dmdb_num_cells") result(n)   ! num_cell, main.F90:390
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: n
    end function
    ! restart path inputinfo.f:76-101 + main.F90:205-321, 389-424; sv is the reference's sv(6,noptotal)
    function dmdb_set_state(handle, replica, sv, bptnr) bind(C, name="dmdb_set_state") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: replica          ! 0-based replica, -1 = all
      real(c_double), intent(in) :: sv(6, *)
      type(c_ptr), value :: bptnr               ! c_loc(bptnr) or c_null_ptr
      integer(c_int) :: rc
    end function
    function dmdb_set_state_all(handle, sv_all, bptnr_all) bind(C, name="dmdb_set_state_all") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: sv_all(*)
      type(c_ptr), value :: bptnr_all
      integer(c_int) :: rc
    end function
    function dmdb_get_state_all(handle, sv_all, bptnr_all) bind(C, name="dmdb_get_state_all") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: sv_all(*)
      type(c_ptr), value :: bptnr_all
      integer(c_int) :: rc
    end function
    function dmdb_set_temperature(handle, replica, tstar) bind(C, name="dmdb_set_temperature") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      real(c_double), value :: tstar
      integer(c_int) :: rc
    end function
    ! nbor()  -- nbor.f:33-137 + cell_add.f:12-28
    function dmdb_nbor(handle) bind(C, name="dmdb_nbor") result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    ! events() -- events.f:23-123
    function dmdb_predict_all(handle) bind(C, name="dmdb_predict_all") result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    ! the main loop main.F90:484-1258: ncoll calendar events per replica
    function dmdb_run(handle, n_events, stats) bind(C, name="dmdb_run") result(rc)
      import :: c_ptr, c_int, c_int64_t, dmdb_stats
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: n_events
      type(dmdb_stats), intent(out) :: stats
      integer(c_int) :: rc
    end function
    ! the same loop, returning right after the next output pseudo-event (main.F90:1191-1246)
    function dmdb_run_until_output(handle, max_events, stats) bind(C, name="dmdb_run_until_output") result(rc)
      import :: c_ptr, c_int, c_int64_t, dmdb_stats
      type(c_ptr), value :: handle
      integer(c_int64_t), value :: max_events
      type(dmdb_stats), intent(out) :: stats
      integer(c_int) :: rc
    end function
    ! main.F90:1288-1295
    function dmdb_sync_positions(handle) bind(C, name="dmdb_sync_positions") result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int) :: rc
    end function
    function dmdb_get_cells(handle, replica, cell_of_bead) bind(C, name="dmdb_get_cells") result(rc)
      import :: c_ptr, c_int, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      integer(c_int32_t), intent(out) :: cell_of_bead(*)
      integer(c_int) :: rc
    end function
    function dmdb_get_nbors(handle, replica, down, offsets, nb) bind(C, name="dmdb_get_nbors") result(rc)
      import :: c_ptr, c_int, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: replica, down
      integer(c_int32_t), intent(out) :: offsets(*)
      type(c_ptr), value :: nb                  ! c_null_ptr to query sizes
      integer(c_int) :: rc
    end function
    ! tim / nptnr / coltype of header.f:20-22,47 (noptotal+3 entries each)
    function dmdb_get_calendar(handle, replica, tim, nptnr, coltype) bind(C, name="dmdb_get_calendar") result(rc)
      import :: c_ptr, c_int, c_int32_t, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      real(c_double), intent(out) :: tim(*)
      integer(c_int32_t), intent(out) :: nptnr(*), coltype(*)
      integer(c_int) :: rc
    end function
    function dmdb_get_state(handle, replica, sv, bptnr, identity, extra_repuls, t, tfalse, coll) &
        bind(C, name="dmdb_get_state") result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      type(c_ptr), value :: sv, bptnr, identity, extra_repuls, t, tfalse, coll   ! c_loc(...) or c_null_ptr
      integer(c_int) :: rc
    end function
    function dmdb_get_evcode(handle, replica, n_pairs, i, j, code) bind(C, name="dmdb_get_evcode") result(rc)
      import :: c_ptr, c_int, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: replica, n_pairs
      integer(c_int32_t), intent(in) :: i(*), j(*)
      integer(c_int32_t), intent(out) :: code(*)
      integer(c_int) :: rc
    end function
    ! energy(ered,tred,sumvel,hb_alpha,hb_ii,hb_ij,ehh_ii,ehh_ij) -- energy.f:25-101
    function dmdb_energy_of(handle, replica, e) bind(C, name="dmdb_energy_of") result(rc)
      import :: c_ptr, c_int, dmdb_energy
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      type(dmdb_energy), intent(out) :: e
      integer(c_int) :: rc
    end function
    function dmdb_get_event_log(handle, replica, first, n, out, n_out) bind(C, name="dmdb_get_event_log") result(rc)
      import :: c_ptr, c_int, c_int64_t, dmdb_event
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      integer(c_int64_t), value :: first, n
      type(dmdb_event), intent(out) :: out(*)
      integer(c_int64_t), intent(out) :: n_out
      integer(c_int) :: rc
    end function
    function dmdb_get_replica_stats(handle, replica, s) bind(C, name="dmdb_get_replica_stats") result(rc)
      import :: c_ptr, c_int, dmdb_stats
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      type(dmdb_stats), intent(out) :: s
      integer(c_int) :: rc
    end function
    function dmdb_get_batch_stats(handle, replica, out) bind(C, name="dmdb_get_batch_stats") result(rc)
      import :: c_ptr, c_int, c_int64_t
      type(c_ptr), value :: handle
      integer(c_int), value :: replica
      integer(c_int64_t), intent(out) :: out(16)
      integer(c_int) :: rc
    end function
    function dmdb_potential_energies(handle, epot, tstar) bind(C, name="dmdb_potential_energies") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: epot(*), tstar(*)
      integer(c_int) :: rc
    end function
    function dmdb_apply_temperatures(handle, tstar_new) bind(C, name="dmdb_apply_temperatures") result(rc)
      import :: c_ptr, c_int, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: tstar_new(*)
      integer(c_int) :: rc
    end function
    ! replica exchange on the device (new functionality; the reference runs temp_0xx one after the other,
    ! qfile/script.sh:11-18).  comm: an ncclComm_t the host created (c_loc / c_ptr), or c_null_ptr for the
    ! communicator of dmdb_comm_init.
    function dmdb_nccl_unique_id(id) bind(C, name="dmdb_nccl_unique_id") result(rc)
      import :: c_int, c_char
      character(kind=c_char), intent(out) :: id(128)
      integer(c_int) :: rc
    end function
    function dmdb_comm_init(handle, id, world, rank) bind(C, name="dmdb_comm_init") result(rc)
      import :: c_ptr, c_int, c_char
      type(c_ptr), value :: handle
      character(kind=c_char), intent(in) :: id(128)
      integer(c_int), value :: world, rank
      integer(c_int) :: rc
    end function
    function dmdb_exchange(handle, comm, step, seed, ladder_size, stats) bind(C, name="dmdb_exchange") result(rc)
      import :: c_ptr, c_int, c_int32_t, c_int64_t, dmdb_exchange_stats
      type(c_ptr), value :: handle, comm
      integer(c_int64_t), value :: step, seed
      integer(c_int32_t), value :: ladder_size
      type(dmdb_exchange_stats), intent(out) :: stats
      integer(c_int) :: rc
    end function
    ! the same decision + temperature change when the host gathers (E_pot, T*) itself (MPI_Allgather):
    ! gathered(2, n_replicas, world), rank-major
    function dmdb_exchange_gathered(handle, gathered, world, rank, step, seed, ladder_size, stats) &
        bind(C, name="dmdb_exchange_gathered") result(rc)
      import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double, dmdb_exchange_stats
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: gathered(*)
      integer(c_int), value :: world, rank
      integer(c_int64_t), value :: step, seed
      integer(c_int32_t), value :: ladder_size
      type(dmdb_exchange_stats), intent(out) :: stats
      integer(c_int) :: rc
    end function
    ! beta-sheet observables of every replica from resident state (results/r/fibril_list_assign.f definitions):
    ! out(8, n_replicas)
    function dmdb_sheet_observables(handle, out) bind(C, name="dmdb_sheet_observables") result(rc)
      import :: c_ptr, c_int, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int32_t), intent(out) :: out(*)
      integer(c_int) :: rc
    end function
    function dmdb_device_fill(device, n_replicas, n_service_ctas) bind(C, name="dmdb_device_fill") result(rc)
      import :: c_int, c_int32_t
      integer(c_int), value :: device
      integer(c_int32_t), intent(out) :: n_replicas, n_service_ctas
      integer(c_int) :: rc
    end function
    function dmdb_set_service_ctas(handle, n) bind(C, name="dmdb_set_service_ctas") result(rc)
      import :: c_ptr, c_int
      type(c_ptr), value :: handle
      integer(c_int), value :: n
      integer(c_int) :: rc
    end function
  end interface

contains

  ! NUL-terminated C string of dmdb_last_error -> Fortran character
  function dmdb_error_message(handle) result(msg)
    type(c_ptr), intent(in) :: handle
    character(len=:), allocatable :: msg
    type(c_ptr) :: p
    character(kind=c_char), pointer :: s(:)
    integer :: n, k
    p = dmdb_last_error(handle)
    if (.not. c_associated(p)) then
      msg = ''
      return
    end if
    call c_f_pointer(p, s, [1024])
    n = 0
    do while (n < 1024)
      if (s(n + 1) == c_null_char) exit
      n = n + 1
    end do
    allocate (character(len=n) :: msg)
    do k = 1, n
      msg(k:k) = s(k)
    end do
  end function

end module dmdb200
