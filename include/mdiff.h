/*
 * libmdiff — C ABI of the B200-native Morphable Diffusion multi-view denoise hot path.
 *
 * The reference (xiyichen/morphablediffusion) has no FFI: its boundary is a set of Python classes resolved by
 * dotted name (ldm/util.py:217-232, configs/facescape.yaml:3,27).  This header is the C boundary a binding of
 * those classes calls; each entry point cites the reference function it replaces.  Conventions:
 *   - every function returns 0 on success, a negative code on error; md_last_error() gives the message;
 *   - all tensor pointers are DEVICE pointers owned by the caller unless stated otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing here takes or returns a torch type.
 */
#ifndef MDIFF_H_
#define MDIFF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MD_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ library-level */
MD_API int md_version(void);
MD_API const char* md_last_error(void);
/* number of kernels this library launched since the last reset (bench.py "gpu_launches") */
MD_API long long md_launch_count(void);
MD_API void md_reset_launch_count(void);

/* ------------------------------------------------------------------ op level: tensor-core implicit GEMM
 * Dense contraction used by every Conv2d/Conv3d/ConvTranspose3d/Linear on the path
 * (reference: torch.nn.Conv2d in ldm/modules/diffusionmodules/openaimodel.py:199-276, nn.Linear in
 * ldm/modules/attention.py:161-203, nn.Conv3d / ConvTranspose3d in ldm/models/diffusion/network.py:285-347).
 *   A   : bf16 channels-last activations [B][D][H][W][Cpitch], the first Cin channels are contracted
 *   Wt  : bf16 packed weights [N][ntaps*Cin] (K-major; tap-major then channel)
 *   out : row-major [rows][ldo]; row = ((b*OD + z*os+op)*OH + y*os+op)*OW + x*os+op
 */
enum { MD_ACT_NONE = 0, MD_ACT_SILU = 1, MD_ACT_RELU = 2, MD_ACT_GEGLU = 3, MD_ACT_GELU = 4 };

typedef struct md_conv_gemm_args {
  const void* A;      /* bf16 */
  int B, D, H, W;     /* input dims (a plain GEMM uses D=H=1, W=rows per sample, B=samples) */
  int Cin, Cpitch;    /* Cpitch = 0 -> Cin */
  const void* Wt;     /* bf16 [N][ntaps*Cin] */
  int N;
  int ntaps;
  int tap[27][3];     /* (dx,dy,dz) added to the output coordinate to address the input */
  int OD, OH, OW;     /* output dims; 0 -> same as input */
  int os[3], op[3];   /* output coordinate = input coordinate * os + op (x,y,z); os 0 -> 1 */
  const float* bias;        /* [N] or NULL */
  const float* rowvec;      /* [B][rowvec_ld] per-sample additive vector or NULL */
  int rowvec_ld;
  const float* res_f32;     /* residual added after activation, [rows][ldo] or NULL */
  const void* res_bf16;
  float* out_f32;           /* either or both outputs */
  void* out_bf16;
  int ldo;                  /* 0 -> N (N/2 for GEGLU) */
  int act;
  float out_scale;          /* 0 -> 1 */
  int BN;                   /* tile N override (64/128/160/256), 0 = auto */
} md_conv_gemm_args;

MD_API int md_op_conv_gemm(const md_conv_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDIFF_H_ */
