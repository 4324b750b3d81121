// Standalone probe: validates the UMMA descriptor conventions the conv kernels rely on.
//   (1) K-major SW128 A operand addressed at an arbitrary ROW shift inside a 1024B-aligned slab
//       (the "one slab, nine shifted taps" trick of the 3x3 implicit GEMM),
//   (2) MN-major SW128 A (two 64-channel atoms LBO apart, i.e. two different row shifts) x MN-major B
//       (the wgrad formulation: contraction over pixels).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../bruno_b200/csrc/ptx.cuh"

using namespace bruno;

struct ProbeArgs {
  uint32_t a_off;      // byte offset of A start inside slab
  uint32_t a_kstep;    // byte advance of A start per UMMA_K step
  uint32_t b_kstep;
  uint32_t nk;         // number of K=16 steps
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_base_off;  // descriptor base_offset field for A
  uint32_t idesc;
  uint32_t slab_bytes, b_bytes;
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* __restrict__ slab_g, const uint8_t* __restrict__ b_g, float* __restrict__ out,
             ProbeArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* slab = smem;                       // 1024-aligned
  uint8_t* bw = smem + 64 * 1024;             // 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 64);
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], a.slab_bytes + a.b_bytes);
    bulk_g2s(slab, slab_g, a.slab_bytes, &bars[0]);
    bulk_g2s(bw, b_g, a.b_bytes, &bars[0]);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    for (uint32_t k = 0; k < a.nk; ++k) {
      uint64_t ad = umma_desc_sw128(smem_u32(slab) + a.a_off + k * a.a_kstep, a.a_lbo, a.a_sbo);
      ad |= static_cast<uint64_t>(a.a_base_off & 7) << 49;
      uint64_t bd = umma_desc_sw128(smem_u32(bw) + k * a.b_kstep, a.b_lbo, a.b_sbo);
      umma_bf16(tmem, ad, bd, a.idesc, k > 0);
    }
    umma_commit(&bars[1]);
  }
  __syncthreads();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + half * 32, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + threadIdx.x % 32) * 64 + half * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = static_cast<uint32_t>(h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// logical [rows][64] bf16 -> SW128 image (16B chunk c of row r stored at chunk c ^ (r & 7))
static void swizzle_rows(const std::vector<uint16_t>& logical, std::vector<uint16_t>& img, int rows) {
  img.assign(static_cast<size_t>(rows) * 64, 0);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 8; ++c)
      for (int e = 0; e < 8; ++e) img[r * 64 + ((c ^ (r & 7)) * 8) + e] = logical[r * 64 + c * 8 + e];
}

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                      \
    }                                                                               \
  } while (0)

int main() {
  const int R = 320;  // slab rows
  std::vector<uint16_t> slab_l(R * 64), slab_img, b_l(64 * 64), b_img;
  srand(1);
  for (auto& x : slab_l) x = f2bf((rand() % 2001 - 1000) / 1000.f);
  for (auto& x : b_l) x = f2bf((rand() % 2001 - 1000) / 1000.f);
  swizzle_rows(slab_l, slab_img, R);
  swizzle_rows(b_l, b_img, 64);
  uint8_t *d_slab, *d_b;
  float* d_out;
  CK(cudaMalloc(&d_slab, R * 128));
  CK(cudaMalloc(&d_b, 64 * 128));
  CK(cudaMalloc(&d_out, 128 * 64 * 4));
  CK(cudaMemcpy(d_slab, slab_img.data(), R * 128, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_b, b_img.data(), 64 * 128, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  std::vector<float> out(128 * 64);
  int fails = 0;

  // ---- test 1: K-major, arbitrary row shift ----
  const int shifts[] = {0, 8, 1, 3, 7, 29, 30, 31, 58, 59, 60, 17};
  for (int bo_mode = 0; bo_mode < 2; ++bo_mode) {
    for (int s : shifts) {
      ProbeArgs a{};
      a.a_off = s * 128;
      a.a_kstep = 32;
      a.b_kstep = 32;
      a.nk = 4;
      a.a_lbo = 16; a.a_sbo = 1024; a.b_lbo = 16; a.b_sbo = 1024;
      a.a_base_off = bo_mode ? (s & 7) : 0;
      a.idesc = umma_idesc_bf16(128, 64, 0, 0);
      a.slab_bytes = R * 128;
      a.b_bytes = 64 * 128;
      CK(cudaMemset(d_out, 0, 128 * 64 * 4));
      probe_kernel<<<1, 128, 100 * 1024>>>(d_slab, d_b, d_out, a);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(out.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          double ref = 0;
          for (int k = 0; k < 64; ++k) ref += double(bf2f(slab_l[(s + m) * 64 + k])) * bf2f(b_l[n * 64 + k]);
          maxerr = fmax(maxerr, fabs(ref - out[m * 64 + n]));
        }
      bool ok = maxerr < 1e-3;
      if (!ok && bo_mode == 0) fails++;
      printf("KMAJOR shift=%3d base_off_mode=%d maxerr=%.3e %s\n", s, bo_mode, maxerr, ok ? "OK" : "FAIL");
    }
  }

  // ---- test 2: MN-major A (two atoms at different row shifts) x MN-major B ----
  // D[m][n] = sum_{kk<16*nk} slab[(m<64?da:db)+kk][m%64] * Y[kk][n];  Y = rows of b_l (64 rows x 64 co)
  const int pairs[][2] = {{0, 8}, {0, 1}, {3, 4}, {1, 30}, {29, 31}, {5, 64}, {2, 2}};
  for (auto& pr : pairs) {
    int da = pr[0], db = pr[1];
    ProbeArgs a{};
    a.a_off = da * 128;
    a.a_kstep = 16 * 128;
    a.b_kstep = 16 * 128;
    a.nk = 4;  // 64 pixels
    a.a_lbo = (db - da) * 128; a.a_sbo = 1024;
    a.b_lbo = 16; a.b_sbo = 1024;
    a.a_base_off = 0;
    a.idesc = umma_idesc_bf16(128, 64, 1, 1);
    a.slab_bytes = R * 128;
    a.b_bytes = 64 * 128;
    CK(cudaMemset(d_out, 0, 128 * 64 * 4));
    probe_kernel<<<1, 128, 100 * 1024>>>(d_slab, d_b, d_out, a);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out.data(), d_out, 128 * 64 * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 64; ++n) {
        double ref = 0;
        int sh = m < 64 ? da : db;
        for (int kk = 0; kk < 64; ++kk) ref += double(bf2f(slab_l[(sh + kk) * 64 + (m % 64)])) * bf2f(b_l[kk * 64 + n]);
        maxerr = fmax(maxerr, fabs(ref - out[m * 64 + n]));
      }
    bool ok = maxerr < 1e-3;
    if (!ok) fails++;
    printf("MNMAJOR da=%2d db=%2d maxerr=%.3e %s\n", da, db, maxerr, ok ? "OK" : "FAIL");
  }
  printf("PROBE %s (%d failing variants)\n", fails ? "FAILED" : "PASSED", fails);
  return 0;
}
