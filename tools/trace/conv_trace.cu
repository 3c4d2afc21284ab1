// Experiment harness (not part of libspvd_b200.so): the f16 tcgen05 sparse convolution compiled with -DSPVD_TRACE so
// every CTA records clock marks of its pipeline phases.  Built and driven by tools/_conv_trace.py on the GPU box.
#define SPVD_TRACE 1
#include "../../spvd_b200/csrc/conv_umma_impl.cuh"

using namespace spvd;

namespace spvd {
thread_local char g_err[512];
void set_error(const char* fmt, ...) { (void)fmt; }
}  // namespace spvd

template <int BN>
static int run(const void* in, const void* wp, const int* map_t, int ld, const unsigned* mask, const int* out_rows, int n_out,
               int kvol, int cin, int cout, const float* bias, float* out, int lite, unsigned long long* trace) {
  cudaMemcpyToSymbol(g_spvd_trace, &trace, sizeof(trace));
  if (lite)
    return launch_umma<BN, 1, true, true, true>((const float*)in, (const float*)wp, map_t, ld, mask, n_out, kvol, cin, cout, bias,
                                                nullptr, out, 1, 0, out_rows);
  return launch_umma<BN, 1, true, false, true>((const float*)in, (const float*)wp, map_t, ld, mask, n_out, kvol, cin, cout, bias,
                                               nullptr, out, 1, 0, out_rows);
}

extern "C" int conv_trace(const void* in, const void* wp, const int* map_t, int ld, const unsigned* mask, const int* out_rows,
                          int n_out, int kvol, int cin, int cout, const float* bias, float* out, int lite,
                          unsigned long long* trace) {
  switch (cout) {
    case 32: return run<32>(in, wp, map_t, ld, mask, out_rows, n_out, kvol, cin, cout, bias, out, lite, trace);
    case 64: return run<64>(in, wp, map_t, ld, mask, out_rows, n_out, kvol, cin, cout, bias, out, lite, trace);
    case 128: return run<128>(in, wp, map_t, ld, mask, out_rows, n_out, kvol, cin, cout, bias, out, lite, trace);
    case 192: return run<192>(in, wp, map_t, ld, mask, out_rows, n_out, kvol, cin, cout, bias, out, lite, trace);
  }
  return -1;
}
