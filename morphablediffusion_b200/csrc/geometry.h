// Launchers of the geometry kernels (geometry.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace md {

// NoisyTargetViewEncoder parameters (device pointers into the caller's fp32 state dict; torch layouts).
struct EncWeightsHost {
  const float* init_w; const float* init_b;
  struct Res {
    const float* te_w; const float* te_b; const float* ve_w; const float* ve_b;
    const float* gn0_w; const float* gn0_b; const float* c0_w; const float* c0_b;
    const float* gn1_w; const float* gn1_b; const float* c1_w; const float* c1_b;
  } res[3];
  const float* fgn_w; const float* fgn_b; const float* fc_w; const float* fc_b;
};

int launch_voxelize(const float* vertices, int nv, int32_t* coord, int32_t* out_sh, float* bounds, cudaStream_t st);
int launch_target_encoder(const float* x, const float* t_embed, const float* v_embed, const EncWeightsHost& w,
                          float* out, int n_views, int tdim, int vdim, int S, cudaStream_t st);
int launch_vertex_features(const float* feats, const float* proj, int ortho, int size, int V, float length,
                           const float* vertices, int nv, int n_views, float* out, cudaStream_t st);
int launch_smpl_scatter(const float* vsum, float inv_views, const float* W, const float* bias,
                        const int32_t* row_vertex, int n_rows, float* out, cudaStream_t st);
int launch_peer_push(const float* vsum, int n_floats, float* const* peer_data, unsigned* const* peer_flags,
                     unsigned* seq_ticket, int rank, int world, size_t slot_floats, cudaStream_t st);
int launch_smpl_scatter_peer(const float* slots, const unsigned* flags, const unsigned* seq_ticket, int world,
                             size_t slot_floats, float inv_views, const float* W, const float* bias,
                             const int32_t* row_vertex, int n_rows, float* out, int* err, cudaStream_t st);
int launch_sparse_conv(const float* in, const int32_t* nbr, const float* W, const float* scale, const float* shift,
                       float* out, int n_rows, int Cin, int Cout, cudaStream_t st);
int launch_volume_resample(const float* feat, const int32_t* idx, const float* wgt, float* vol, int npts,
                           cudaStream_t st);
int launch_frustum_points(const float* cam, int ortho, int D, int size, float length, float frustum_len, float* pts,
                          int n_views, cudaStream_t st);
// batch construction (generate_face.py:203-249)
int launch_affine_points(const float* v, int n, const float* A9_host, const float* b3_host, float* out, cudaStream_t st);
int launch_images_to_u8(const float* img, uint8_t* out, int n, int HW, cudaStream_t st);
int launch_frustum_gather(const float* vol, const float* pts, int V, void* out_bf16, size_t npts, cudaStream_t st);

}  // namespace md
