// dmd_cuda.cu -- libdmdb200.so: CUDA backend (sm_100a) of the C ABI in include/dmdb200.h.
//
// Kernels (one warp per replica; 4 replicas per CTA; the 28x28 pair tables are staged in shared memory):
//   dmd_event_loop_kernel    the persistent event loop, main.F90:484-1258 (dmdb_run)
//   dmd_start_kernel         run start: first ghost time (main.F90:408-416) + nbor() + events()
//   dmd_nbor_kernel          cell_add.f + nbor.f  (dmdb_nbor)
//   dmd_predict_all_kernel   events.f             (dmdb_predict_all)
//   dmd_sync_positions_kernel main.F90:1288-1295
//   dmd_energy_kernel        energy.f
//   dmd_retemp_kernel        replica-exchange temperature change on resident state (new functionality)
//   dmd_evcode_kernel        ev_code(i,j) read-back for the parity tests
// There is no CPU fallback: be::init fails when no CUDA device is present.
#include <cuda_runtime.h>

#include <stdexcept>
#include <string>

#include "dmd_engine.h"
#include "dmd_types.h"

#define CUDA_OK(x)                                                                                   \
  do {                                                                                               \
    cudaError_t e_ = (x);                                                                            \
    if (e_ != cudaSuccess) throw std::runtime_error(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace dmd {

constexpr int WARPS_PER_CTA = 4;
#ifndef DMD_MIN_CTAS
#define DMD_MIN_CTAS 4  // 4 CTAs x 4 warps per SM at <= 128 registers per thread
#endif

__device__ __forceinline__ const PairTables* stage_tables(const DevArrays& d, PairTables* smem) {
  const double* src = reinterpret_cast<const double*>(d.tables);
  double* dst = reinterpret_cast<double*>(smem);
  for (int k = threadIdx.x; k < (int)(sizeof(PairTables) / 8); k += blockDim.x) dst[k] = src[k];
  __syncthreads();
  return smem;
}

__device__ __forceinline__ int32_t* warp_queue() {
  __shared__ int32_t s_cq[WARPS_PER_CTA][CQ_CAP];
  return s_cq[threadIdx.x >> 5];
}

__device__ __forceinline__ int replica_of_warp(int r0, int nrep) {
  int w = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
  return w < nrep ? r0 + w : -1;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, DMD_MIN_CTAS) dmd_event_loop_kernel(DevArrays d, int r0, int nrep, long long n_events) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  if (r.error == 0) run_events(r, n_events);
  rep_save(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_start_kernel(DevArrays d, int r0, int nrep) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  if (d.sys->canon) {  // main.F90:408-416
    double tgho = 0.0;
    while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform(r.seed, r.ctr);
    if (Warp::lane() == 0) r.cal[r.N].t = -1.0 * dmd_log(tgho) * r.avegtime * .0000001;
  }
  nbor(r);
  predict_all(r);
  rep_save(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_nbor_kernel(DevArrays d, int r0, int nrep) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  nbor(r);
  rep_save(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_predict_all_kernel(DevArrays d, int r0, int nrep) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  predict_all(r);
  rep_save(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_sync_positions_kernel(DevArrays d, int r0, int nrep) {
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, d.tables, warp_queue(), rid);
  sync_positions(r);
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_energy_kernel(DevArrays d, int r0, int nrep, OutRec* eout) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  OutRec o;
  energy_of(r, o);
  if (Warp::lane() == 0) eout[rid] = o;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) dmd_retemp_kernel(DevArrays d, int r0, int nrep, const double* tstar_new) {
  __shared__ PairTables stab;
  const PairTables* tab = stage_tables(d, &stab);
  int rid = replica_of_warp(r0, nrep);
  if (rid < 0) return;
  const double tn = tstar_new[rid];
  if (!(tn > 0.0)) return;  // <= 0: leave this replica untouched
  Rep r;
  rep_bind(r, d, tab, warp_queue(), rid);
  retemp(r, tn);
  rep_save(r);
}

__global__ void dmd_evcode_kernel(DevArrays d, int rid, int n_pairs, int32_t* buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_pairs) return;
  const int N = d.sys->N;
  int i = buf[k] - 1, j = buf[n_pairs + k] - 1;
  const BeadRec* rec = d.rec + (size_t)rid * N;
  int sc = static_code(*d.sys, d.meta[i], d.chain[i], i, d.meta[j], d.chain[j], j);
  buf[2 * n_pairs + k] = overlay_code(sc, i, rec[i], j, rec[j]);
}

}  // namespace dmd

namespace be {

static cudaStream_t g_stream = nullptr;
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

inline bool init(int device, std::string& err) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    err = std::string("libdmdb200 needs a CUDA device (sm_100a); none found: ") + cudaGetErrorString(e);
    return false;
  }
  if (device < 0 || device >= n) {
    err = "CUDA device ordinal out of range";
    return false;
  }
  if ((e = cudaSetDevice(device)) != cudaSuccess) {
    err = cudaGetErrorString(e);
    return false;
  }
  if (!g_stream) {
    cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking);
    cudaEventCreate(&g_ev0);
    cudaEventCreate(&g_ev1);
  }
  return true;
}
inline void* alloc(size_t n) {
  void* p = nullptr;
  CUDA_OK(cudaMalloc(&p, n ? n : 1));
  return p;
}
inline void release(void* p) { cudaFree(p); }
inline void h2d(void* d, const void* h, size_t n) {
  CUDA_OK(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
}
inline void d2h(void* h, const void* d, size_t n) {
  CUDA_OK(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
}
inline void zero(void* d, size_t n) { CUDA_OK(cudaMemsetAsync(d, 0, n, g_stream)); }
inline void fill_i32(int32_t* d, int v, size_t n) {
  if (v == -1) CUDA_OK(cudaMemsetAsync(d, 0xff, n * 4, g_stream));
  else if (v == 0) CUDA_OK(cudaMemsetAsync(d, 0, n * 4, g_stream));
  else throw std::runtime_error("fill_i32: unsupported value");
}

// one launcher per device operation; times the kernel with CUDA events on the launching stream
inline void run_op(const dmd::DevArrays& d, int op, int r0, int nrep, long long arg, int32_t* ibuf, dmd::OutRec* eout,
                   double* ms, int* launches) {
  using namespace dmd;
  const int block = WARPS_PER_CTA * 32;
  const int grid = (nrep + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
  CUDA_OK(cudaEventRecord(g_ev0, g_stream));
  switch (op) {
    case 0: dmd_start_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep); break;
    case 1: dmd_nbor_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep); break;
    case 2: dmd_predict_all_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep); break;
    case 3: dmd_event_loop_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep, arg); break;
    case 4: dmd_sync_positions_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep); break;
    case 5: dmd_energy_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep, eout); break;
    case 6: dmd_evcode_kernel<<<((int)arg + 127) / 128, 128, 0, g_stream>>>(d, r0, (int)arg, ibuf); break;
    case 7: dmd_retemp_kernel<<<grid, block, 0, g_stream>>>(d, r0, nrep, (const double*)ibuf); break;
    default: throw std::runtime_error("unknown device op");
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(g_ev1, g_stream));
  CUDA_OK(cudaStreamSynchronize(g_stream));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, g_ev0, g_ev1));
  if (ms) *ms = t;
  if (launches) *launches = 1;
}

}  // namespace be

#include "dmd_capi_impl.h"
