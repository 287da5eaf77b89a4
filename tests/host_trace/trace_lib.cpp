// trace_lib.cpp -- TEST SCAFFOLDING, never shipped and never loaded by the product path.
//
// Compiles the engine source (csrc/dmd_engine.h) with DMD_HOST_TRACE: a 1-lane "warp" executed on the CPU,
// behind the same C ABI, so that the event-loop LOGIC (calendar, cascades, bookkeeping, rebuilds) can be
// diffed against the oracle in the GPU-less build container (tests/test_engine_logic_hosttrace.py).
// It exercises none of the 32-lane collectives; the real parity tests are the `-m gpu` ones.
#define DMD_HOST_TRACE 1
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_engine.h"
#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_types.h"

namespace be {
inline bool init(int, std::string&) { return true; }
inline void* alloc(size_t n) { return std::calloc(n ? n : 1, 1); }
inline void release(void* p) { std::free(p); }
inline void h2d(void* d, const void* h, size_t n) { std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n) { std::memcpy(h, d, n); }
inline void zero(void* d, size_t n) { std::memset(d, 0, n); }
inline void fill_i32(int32_t* d, int v, size_t n) {
  for (size_t k = 0; k < n; k++) d[k] = v;
}
inline void run_op(const dmd::DevArrays& d, int op, int r0, int nrep, long long arg, int32_t* ibuf, dmd::OutRec* eout,
                   double* ms, int* launches) {
  using namespace dmd;
  if (op == 6) {
    const int N = d.sys->N, n_pairs = (int)arg;
    const BeadRec* rec = d.rec + (size_t)r0 * N;
    for (int k = 0; k < n_pairs; k++) {
      int i = ibuf[k] - 1, j = ibuf[n_pairs + k] - 1;
      int sc = static_code(*d.sys, d.meta[i], d.chain[i], i, d.meta[j], d.chain[j], j);
      ibuf[2 * n_pairs + k] = overlay_code(sc, i, rec[i], j, rec[j]);
    }
  } else {
    for (int rid = r0; rid < r0 + nrep; rid++) {
      Rep r;
      int32_t cq[CQ_CAP];
      rep_bind(r, d, d.tables, cq, rid);
      switch (op) {
        case 0:
          if (d.sys->canon) {
            double tgho = 0.0;
            while (tgho < 1e-18 || tgho == 1.0) tgho = rng_uniform(r.seed, r.ctr);
            r.cal[r.N].t = -1.0 * dmd_log(tgho) * r.avegtime * .0000001;
          }
          nbor(r);
          predict_all(r);
          rep_save(r);
          break;
        case 1: nbor(r); rep_save(r); break;
        case 2: predict_all(r); rep_save(r); break;
        case 3: if (r.error == 0) run_events(r, arg); rep_save(r); break;
        case 4: sync_positions(r); break;
        case 5: { OutRec o; energy_of(r, o); eout[rid] = o; } break;
        case 7: { double tn = ((const double*)ibuf)[rid]; if (tn > 0.0) { retemp(r, tn); rep_save(r); } } break;
        default: throw std::runtime_error("unknown op");
      }
    }
  }
  if (ms) *ms = 0;
  if (launches) *launches = 0;
}
}  // namespace be

#include "../../parallel_dmd_for_biomolecules_b200/csrc/dmd_capi_impl.h"
