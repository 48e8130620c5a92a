// Host-side view of the tensor-core sweep's geometry and layout functions (reevr_b200/csrc/kernels_tc.cuh), for
// tests/test_tc_layout.py: the same inline functions the kernels use, compiled by g++ (no GPU needed).
#include "../../reevr_b200/csrc/kernels_tc.cuh"

extern "C" {
void tc_geom(int P, int nb, int* out) {
  const pc::tc::Geom g = pc::tc::make_geom(P, nb);
  out[0] = g.Q; out[1] = g.nchunk; out[2] = g.nseg; out[3] = g.ntile; out[4] = g.rows;
}
int tc_geom_ok(int P, int nb, int B) { return pc::tc::geom_ok(pc::tc::make_geom(P, nb), B) ? 1 : 0; }
unsigned tc_sw128(unsigned r, unsigned e) { return pc::tc::sw128(r, e); }
unsigned long long tc_xt_index(long long line, int pl, int e, long long R, int jj, int rows) { return pc::tc::xt_index(line, pl, e, R, jj, rows); }
void tc_consts(int* out) {
  out[0] = pc::tc::kR; out[1] = pc::tc::kN; out[2] = pc::tc::kStripRows; out[3] = pc::tc::kStripBytes; out[4] = pc::tc::kATileBytes;
  out[5] = pc::tc::kMaxChunks; out[6] = pc::tc::kFlush; out[7] = pc::tc::kSmemBytes;
}
}
