// Host-side orchestration of one CouplingLayerConv (nn_extra_nvp.py:58-187): the whole layer -- c1, R residual
// blocks of two tensor-core convolutions, the two heads, the coupling and the log-det -- is ONE call through the C ABI
// (19 kernel launches forward for R = 8), so the Python host adds one ctypes call per layer, not one per op.
#include "common.cuh"

using namespace bruno;

extern "C" {

size_t bruno_coupling_conv_slabs(int R, int save) { return save ? static_cast<size_t>(2 * R + 1) : 3; }

int bruno_coupling_conv_fwd(const float* x, const float* ld_in, float* y, float* ld_out, int N, int H, int W, int C, int R,
                            int mask_type, int inverse, const float* V1, const float* g1, const float* b1,
                            const void* wfwd, const float* br, int bias_img_stride, const float* Ks, const float* bs,
                            const float* Kt, const float* bt, const float* label_s, const float* label_t, void* slabs,
                            int n_slabs, int c1_ks, void* stream) {
  BRUNO_REQUIRE(x && ld_in && y && ld_out && V1 && g1 && b1 && wfwd && br && Ks && bs && Kt && bt && slabs,
                "bruno_coupling_conv_fwd: null pointer");
  BRUNO_REQUIRE(R >= 1 && (n_slabs >= 3), "bruno_coupling_conv_fwd: need >= 3 slabs (got %d)", n_slabs);
  const bool save = n_slabs >= 2 * R + 1;
  const size_t sb = bruno_slab_bytes(N, H, W), wb = bruno_conv_wpack_bytes();
  auto S = [&](int i) { return static_cast<void*>(static_cast<uint8_t*>(slabs) + static_cast<size_t>(i) * sb); };
  auto Wp = [&](int l) { return static_cast<const void*>(static_cast<const uint8_t*>(wfwd) + static_cast<size_t>(l) * wb); };
  int rc;
  // bias_img_stride = 0: b1 [64], br [2R,64];  != 0: label-conditioned per-image tables, image n of layer l at
  // br + l*64 + n*bias_img_stride (the 2R+1 tables are the columns of one [N, (2R+1)*64] matrix)
  const size_t bl = 64;
  if ((rc = bruno_c1_fwd(x, V1, g1, b1, bias_img_stride, S(0), 0, N, H, W, C, mask_type, c1_ks, stream))) return rc;
  int hi = 0;
  for (int r = 0; r < R; ++r) {
    const int ui = save ? 2 * r + 1 : (hi + 1) % 3;
    const int ni = save ? 2 * r + 2 : (hi + 2) % 3;
    if ((rc = bruno_conv3x3(0, S(hi), Wp(2 * r), br + (2 * r) * bl, bias_img_stride, nullptr, nullptr, S(ui), N, H, W, stream))) return rc;
    if ((rc = bruno_conv3x3(1, S(ui), Wp(2 * r + 1), br + (2 * r + 1) * bl, bias_img_stride, S(hi), nullptr, S(ni), N, H, W, stream))) return rc;
    hi = ni;
  }
  return bruno_heads_coupling(x, S(hi), Ks, bs, Kt, bt, label_s, label_t, ld_in, y, ld_out, 0, N, H, W, C, mask_type, inverse, stream);
}

int bruno_coupling_conv_bwd(const float* x, const float* dy, const float* dld, int N, int H, int W, int C, int R,
                            int mask_type, const float* V1, const float* g1, const float* Vr, const float* gr,
                            const void* wbwd, const float* Ks, const float* bs, const float* Kt, const float* bt,
                            const void* saved_slabs, void* gslabs, float* scratch, float* dx, float* dV1, float* dg1, float* db1,
                            float* dVr, float* dgr, float* dbr, float* dKs, float* dbs, float* dKt, float* dbt, void* stream) {
  return bruno_coupling_conv_bwd_cond(x, dy, dld, N, H, W, C, R, mask_type, V1, g1, Vr, gr, wbwd, Ks, bs, Kt, bt, saved_slabs,
                                      gslabs, scratch, dx, dV1, dg1, db1, dVr, dgr, dbr, dKs, dbs, dKt, dbt, nullptr, nullptr, 0,
                                      nullptr, nullptr, 1, stream);
}

int bruno_coupling_conv_bwd_cond(const float* x, const float* dy, const float* dld, int N, int H, int W, int C, int R,
                                 int mask_type, const float* V1, const float* g1, const float* Vr, const float* gr,
                                 const void* wbwd, const float* Ks, const float* bs, const float* Kt, const float* bt,
                                 const void* saved_slabs, void* gslabs, float* scratch, float* dx, float* dV1, float* dg1,
                                 float* db1, float* dVr, float* dgr, float* dbr, float* dKs, float* dbs, float* dKt, float* dbt,
                                 const float* label_s, float* dtab, int dtab_stride, float* dlabel_s, float* dlabel_t,
                                 int c1_ks, void* stream) {
  BRUNO_REQUIRE(x && dy && dld && V1 && g1 && Vr && gr && wbwd && Ks && bs && Kt && bt && saved_slabs && gslabs && scratch && dx && dV1 &&
                    dg1 && db1 && dVr && dgr && dbr && dKs && dbs && dKt && dbt, "bruno_coupling_conv_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sb = bruno_slab_bytes(N, H, W), wb = bruno_conv_wpack_bytes();
  auto S = [&](int i) { return static_cast<const void*>(static_cast<const uint8_t*>(saved_slabs) + static_cast<size_t>(i) * sb); };
  auto Wp = [&](int l) { return static_cast<const void*>(static_cast<const uint8_t*>(wbwd) + static_cast<size_t>(l) * wb); };
  // GS(0) = gradient w.r.t. the c1 pre-activation; GS(l + 1) = gradient w.r.t. the pre-activation of conv layer l
  auto GS = [&](int i) { return static_cast<void*>(static_cast<uint8_t*>(gslabs) + static_cast<size_t>(i) * sb); };
  cudaMemsetAsync(dVr, 0, static_cast<size_t>(2 * R) * 9 * 64 * 64 * sizeof(float), st);
  cudaMemsetAsync(dbr, 0, static_cast<size_t>(2 * R) * 64 * sizeof(float), st);
  cudaMemsetAsync(dV1, 0, static_cast<size_t>(c1_ks * c1_ks * C) * 64 * sizeof(float), st);
  cudaMemsetAsync(db1, 0, 64 * sizeof(float), st);
  int rc;
  if ((rc = bruno_heads_coupling_bwd_cond(x, S(2 * R), Ks, bs, Kt, bt, dy, dld, dx, GS(2 * R), dKs, dbs, dKt, dbt, scratch, N, H,
                                          W, C, mask_type, label_s, dlabel_s, dlabel_t, stream))) return rc;
  for (int r = R - 1; r >= 0; --r) {
    // g2 = dgrad_c3(g3) * ELU'(u_r);   g3' (or g1) = (dgrad_c2(g2) + g3) * ELU'(block input)
    if ((rc = bruno_conv3x3(2, GS(2 * r + 2), Wp(2 * r + 1), nullptr, 0, S(2 * r + 1), nullptr, GS(2 * r + 1), N, H, W, stream))) return rc;
    if ((rc = bruno_conv3x3(3, GS(2 * r + 1), Wp(2 * r), nullptr, 0, S(2 * r), GS(2 * r + 2), GS(2 * r), N, H, W, stream))) return rc;
  }
  // all 2R weight (and bias) gradients in one launch: conv layer l reads its input slab S(l) and gradient GS(l + 1)
  if ((rc = bruno_conv3x3_wgrad_batched(S(0), GS(1), dVr, dbr, 2 * R, N, H, W, stream))) return rc;
  if ((rc = bruno_c1_bwd(x, V1, g1, GS(0), dx, dV1, db1, N, H, W, C, mask_type, c1_ks, stream))) return rc;
  // label-conditioned layers: gradient of the per-image bias tables = per-image channel sums of the 2R+1 gradient slabs
  if (dtab && (rc = bruno_slab_image_colsum(gslabs, 2 * R + 1, dtab, dtab_stride, N, H, W, stream))) return rc;
  // weight-norm backward, in place over the raw weight gradients
  if ((rc = bruno_conv_wn_bwd(Vr, gr, dVr, dVr, dgr, 2 * R, 9 * 64, stream))) return rc;
  return bruno_conv_wn_bwd(V1, g1, dV1, dV1, dg1, 1, c1_ks * c1_ks * C, stream);
}

}  // extern "C"
