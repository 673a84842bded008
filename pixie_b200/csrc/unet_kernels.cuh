// Launchers for the memory-bound U-Net kernels (see unet_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>

namespace pixie {

enum { kNormNone = 0, kNormLN = 1, kNormGN = 2 };
enum { kActNone = 0, kActLeaky = 1, kActSiLU = 2 };

struct NormArgs {
    const float* x = nullptr;     // fp32 [NB][V][C]
    int V = 0, C = 0;
    const double* stats = nullptr;  // [NB][C][2] = (sum, sum of squares) over V
    int mode = kNormNone;
    int groups = 1;               // GN only
    const float* gamma = nullptr; // LN: [V]; GN: [C]
    const float* beta = nullptr;
    float eps = 1e-5f;
    int act = kActNone;
    __half* dst = nullptr;        // fp16 [NB][V][dst_ld], written at channel offset dst_c0 (may be null)
    __half* dst_lo = nullptr;     // split-precision modes: lo_mode 0 = fp16(y - float(fp16(y))), same layout as dst;
                                  // lo_mode 1 = E5M2 correction operands (see store_hi_lo in unet_kernels.cu)
    int lo_mode = 0;
    int dst_ld = 0, dst_c0 = 0;
    __half* raw_dst = nullptr;    // optional un-normalised fp16 copy of x
    __half* raw_lo = nullptr;
    int raw_ld = 0, raw_c0 = 0;
    int vox_per_block = 0;        // filled by the launcher
};

int launch_moments(const float* x, int NB, int V, int C, double* stats, cudaStream_t st);
int launch_norm_act(NormArgs a, int NB, cudaStream_t st);
int launch_upsample2(const float* x, __half* y, __half* ylo, int lo_mode, int NB, int sp, int C, cudaStream_t st);
int launch_attention(const float* qkv, __half* out, __half* out_lo, int lo_mode, int NB, int T, int C, cudaStream_t st);
int launch_pack_predictions(const float* seg, const float* cont, float* out, int NB, long long V, int n_classes, cudaStream_t st);
int launch_ncdhw_to_ndhwc_f16(const float* x, __half* y, int NB, int C, int Cpad, long long V, cudaStream_t st);

}  // namespace pixie
