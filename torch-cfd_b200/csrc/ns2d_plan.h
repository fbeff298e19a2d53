// Internal glue between the per-(precision, N) kernel translation units and ns2d_api.cu.
#pragma once
enum { TCFD_K_ROWS_INV = 0, TCFD_K_ROWS_FULL = 1, TCFD_K_ROWS_FWD = 2, TCFD_K_COLS = 3 };
typedef struct {
  int n, prec, yt;
  int (*launch)(int which, const void* params, int num_sms, void* stream);
} tcfd_ns2d_entry_t;
