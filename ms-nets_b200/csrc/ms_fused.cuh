// ms_fused.cuh -- interface of the fused MS-volume kernel (ms_fused.cu).
#pragma once
#include "common.cuh"

namespace msn {

// Optional WTA by-product of launch_ms_fused: [subs][N][4][h][w] planes (argmin over each of channels 0-3 with
// np.argmin's tie rule, smallest and second smallest value); all three or none.
struct FusedWta {
  int32_t* idx;
  float *min1, *min2;
};

// true when (windows, disparity count) match what the fused kernel is specialised for.  With
// p->lr the kernel still produces the LEFT view only, laid out for a 16-channel tensor (phase A
// form); slab.cu's launch_slab_right_view derives channels 8-15 from the parked costs.
bool fused_supported(const msn_ms_params* p, int Dn);
size_t fused_workspace_bytes(int N, int H, int W, int Dn, const msn_ms_params* p);
// d_left/d_right: [N][H][W] uint8; out: [N][C][D][h][w]; workspace 256-byte aligned.
// d_mins == nullptr: the whole volume.  d_mins != nullptr: phase A of the disparity-slab path
// (channels 0-3 final, raw costs parked in channels 4-7, per-pixel slab minima in d_mins
// [N][4 or 8][h][w]).
// out_D / out_d0 / accumulate: see ms_fused.cu (slabs of one volume processed on one GPU).
int launch_ms_fused(const uint8_t* d_left, const uint8_t* d_right, int N, int H, int W, const msn_ms_params* p,
                    float* d_out, float* d_mins, char* workspace, cudaStream_t s, int out_D = 0, int out_d0 = 0,
                    int accumulate = 0, const msn_slab_exchange* xchg = nullptr, const FusedWta* wta = nullptr,
                    bool out_bf16 = false);   // out_bf16: d_out holds __nv_bfloat16 (same shape), one-pass path only
// xchg != nullptr: this rank's disparity slab [p->d_begin, +p->d_count) with the AML minimum / denominator
// traded with the other ranks INSIDE the kernel through their peer-mapped exchange tables (no d_mins, no
// phases B/C).  Bytes of one rank's table:
size_t fused_exchange_bytes(int N, int H, int W, const msn_ms_params* p, int sources);   // sources = ranks x sub-slabs
constexpr int kFusedSlabD = 192;   // slab size when a volume above the fused kernel's limit is cut into slabs

int profile_enable(int on);
int profile_read(double* prep_ms, double* sadsob_ms, double* fused_ms, int* calls);

}  // namespace msn
