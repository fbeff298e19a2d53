// Internal glue between the per-(precision, N) kernel translation units and ns2d_api.cu.
#pragma once
// ROWS_EVAL: forward rows + F / residual epilogue (explicit_terms, residual); timed kinds are 0..3
enum { TCFD_K_ROWS_INV = 0, TCFD_K_ROWS_FULL = 1, TCFD_K_ROWS_FWD = 2, TCFD_K_COLS = 3, TCFD_K_ROWS_EVAL = 4 };
// L2 access-policy window of the dataflow launch: the W-slot workspace (H, advt, unit-layout state) is
// marked persisting so that the caller's streaming arrays cannot displace it (bytes = 0: no window)
typedef struct {
  void* base;
  size_t bytes;
  float hit_ratio;
  int grouped;  // 1: grouped items (3 double rows / 4 quads per ticket), 0: single-unit items
} tcfd_flow_window_t;
typedef struct {
  int n, prec, yt;
  int v2;  // 1: ns2d_v2.cuh kernels (packed layouts H2/advt2, TMA tile maps), 0: ns2d_kernels.cuh
  // maps: const tcfd::TileMaps* (v2 cols kernel only, may be null otherwise)
  int (*launch)(int which, const void* params, const void* maps, int num_sms, void* stream);
  // third-generation persistent dataflow kernel (ns2d_flow.cuh): params = const tcfd::FlowParams<T>*;
  // null when the size has no such kernel
  int (*launch_flow)(const void* flow_params, const void* maps, int num_sms, void* stream, const tcfd_flow_window_t* win);
  int flow_ctas_per_sm;  // CTAs per SM the flow kernel's shared memory allows (0: does not fit)
  // whole-call resident kernel for small grids (ns2d_small.cuh, n <= 64): params = const tcfd::FlowParams<T>*, one CTA
  // per sample; null when the size has no such kernel
  int (*launch_small)(const void* flow_params, int batch, void* stream);
} tcfd_ns2d_entry_t;
