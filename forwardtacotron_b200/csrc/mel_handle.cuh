// The handle behind ftb_mel_*: tables of the STFT->mel kernel (stft_mel.cu) and of the inverse path -- mel -> linear
// spectrogram (NNLS) -> Griffin-Lim (griffin_lim.cu).  Built once per (config, device) by ftb_mel_create.
#pragma once
#include <vector>

#include "common.cuh"

namespace ftb {

struct MelTables {  // device pointers
  const float2* window;  // [512]  periodic Hann, as (w[2m], w[2m+1])
  const float2* tw8;     // [8][9]  pass-2 twiddles exp(-2 pi i r k / 64) at [k * 9 + r]
  const float2* tw64;    // [7][64] pass-3 twiddles exp(-2 pi i r k / 512) at [(r - 1) * 64 + k]
  const float2* w1024;   // [513]  exp(-2 pi i k / 1024)
  const float* vr_w;     // [VL][nvrp] tap weights of the virtual rows (0 where a row has fewer taps)
  const int* vr_start;   // [nvrp] first bin of the virtual row (start + VL <= 513)
  const int* row_first;  // [n_mels] first virtual row of the mel row
  const int* row_cnt;    // [n_mels] number of virtual rows
  int n_mels, nvrp, hop;
};

struct MelInverseTables {  // mel filterbank A (n_mels x 513) as sparse rows and columns, its pseudo-inverse, 1 / ||A||^2
  const int* row_ptr;      // [n_mels + 1]
  const int* row_col;      // [nnz]
  const float* row_val;    // [nnz]
  const int* col_ptr;      // [514]
  const int* col_row;      // [nnz]
  const float* col_val;    // [nnz]
  const float* pinv;       // [513][n_mels]  A^T (A A^T)^-1
  float inv_lipschitz;     // 1 / lambda_max(A A^T)
  int n_mels;
};

}  // namespace ftb

struct ftb_mel_handle {
  ftb_mel_config cfg;
  int device = 0;
  std::vector<void*> owned;
  std::vector<float> fb_host;  // dense (n_mels, 513)
  int smem = 0;                // dynamic shared memory of the STFT->mel kernel (tables + per-warp buffers)
  ftb::MelTables tb;
  ftb::MelInverseTables inv;
  ~ftb_mel_handle() {
    for (void* p : owned) cudaFree(p);
  }
};
