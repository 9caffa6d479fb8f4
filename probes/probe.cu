// Hardware probes (test infrastructure, not part of the product ABI).
// They answer layout questions the conv kernels depend on, on a real B200:
//   * how tcgen05.mma reads swizzled / interleaved shared-memory operands when the view starts at a
//     row that is not aligned to the swizzle atom (shifted "tap" views of one halo'd tile),
//   * what a 5-D TMA box load with negative (out-of-bounds) coordinates leaves in shared memory.
#include <stdio.h>
#include <string.h>
#include "../general-medical-image-segmentation-cnn-framework_b200/csrc/ptx.cuh"

using namespace b200;

struct UmmaProbeArgs {
  uint32_t img_bytes;
  uint32_t a_off, a_lbo, a_sbo, a_layout, a_base_mode;
  uint32_t b_off, b_lbo, b_sbo, b_layout, b_base_mode;
  uint32_t idesc;
  uint32_t nsteps, a_step, b_step;
  uint32_t N;
};

__global__ void __launch_bounds__(128, 1) probe_umma_kernel(const uint8_t* __restrict__ img, UmmaProbeArgs p,
                                                            float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (uint32_t i = tid * 16; i < p.img_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(img + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;

  if (tid == 0) {
    const uint32_t sbase = smem_u32(smem);
    for (uint32_t s = 0; s < p.nsteps; ++s) {
      const uint32_t a_addr = sbase + p.a_off + s * p.a_step;
      const uint32_t b_addr = sbase + p.b_off + s * p.b_step;
      const uint32_t abo = p.a_base_mode ? ((a_addr >> 7) & 7) : 0;
      const uint32_t bbo = p.b_base_mode ? ((b_addr >> 7) & 7) : 0;
      const uint64_t ad = make_smem_desc(a_addr, p.a_lbo, p.a_sbo, p.a_layout, abo);
      const uint64_t bd = make_smem_desc(b_addr, p.b_lbo, p.b_sbo, p.b_layout, bbo);
      umma_f16(tbase, ad, bd, p.idesc, s > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (uint32_t c = 0; c < p.N; c += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tbase + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < p.N) out[(warp * 32 + lane) * p.N + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

__global__ void __launch_bounds__(128, 1)
    probe_tma_kernel(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, int c3, int c4,
                     uint32_t box_bytes, uint32_t dump_bytes, uint8_t* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (uint32_t i = tid; i < dump_bytes; i += blockDim.x) smem[i] = 0xAB;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar, box_bytes);
    tma_load_5d(smem, &tmap, &bar, c0, c1, c2, c3, c4);
  }
  mbar_wait(&bar, 0);
  for (uint32_t i = tid; i < dump_bytes; i += blockDim.x) out[i] = smem[i];
}

// TMA bandwidth probe: every CTA streams `iters` boxes through a 4-deep ring.
__global__ void __launch_bounds__(128, 1)
    probe_tma_bw_kernel(const __grid_constant__ CUtensorMap tmap, uint32_t box_bytes, int n1, int n2, int n3,
                        int s1, int s2, int s3, int iters, uint32_t* __restrict__ sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bars[4];
  const uint32_t stage_bytes = (box_bytes + 1023) & ~1023u;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  __syncthreads();
  if (tid == 0) {
    uint32_t acc = 0;
    const int total = n1 * n2 * n3;
    for (int it = 0; it < iters + 4; ++it) {
      if (it >= 4) {
        const int st = (it - 4) & 3;
        mbar_wait(&bars[st], ((it - 4) >> 2) & 1);
        acc += *reinterpret_cast<volatile uint32_t*>(smem + st * stage_bytes);
      }
      if (it < iters) {
        const int st = it & 3;
        int t = (blockIdx.x * iters + it) % total;
        const int i1 = t % n1;
        t /= n1;
        const int i2 = t % n2;
        t /= n2;
        const int i3 = t % n3;
        mbar_arrive_expect_tx(&bars[st], box_bytes);
        tma_load_5d(smem + st * stage_bytes, &tmap, &bars[st], 0, i1 * s1 - 1, i2 * s2 - 1, i3 * s3, 0);
      }
    }
    sink[blockIdx.x] = acc;
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return -1;                                                               \
    }                                                                          \
  } while (0)

extern "C" int probe_umma(const uint8_t* img_host, const UmmaProbeArgs* args, float* out_host) {
  UmmaProbeArgs p = *args;
  uint8_t* d_img;
  float* d_out;
  const uint32_t padded = (p.img_bytes + 15) & ~15u;
  p.img_bytes = padded;
  CK(cudaMalloc(&d_img, padded));
  CK(cudaMemset(d_img, 0, padded));
  CK(cudaMemcpy(d_img, img_host, args->img_bytes, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, 128 * p.N * sizeof(float)));
  CK(cudaMemset(d_out, 0xFF, 128 * p.N * sizeof(float)));
  const int smem = padded + 2048;
  CK(cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_umma_kernel<<<1, 128, smem>>>(d_img, p, d_out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out_host, d_out, 128 * p.N * sizeof(float), cudaMemcpyDeviceToHost));
  cudaFree(d_img);
  cudaFree(d_out);
  return 0;
}

static int encode_map(CUtensorMap* tm, void* dptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return -2;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, dptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   static_cast<CUtensorMapSwizzle>(swizzle), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return -3;
  }
  return 0;
}

// dims/strides/box: 5 entries, innermost first. strides_bytes[i] = byte stride of dim i+1.
extern "C" int probe_tma(const uint8_t* data_host, uint64_t data_bytes, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, int swizzle, const int* coords,
                         uint32_t dump_bytes, uint8_t* out_host) {
  uint8_t *d_data, *d_out;
  CK(cudaMalloc(&d_data, data_bytes));
  CK(cudaMemcpy(d_data, data_host, data_bytes, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, dump_bytes));
  CUtensorMap tm;
  int rc = encode_map(&tm, d_data, 5, dims, strides_bytes, box, swizzle);
  if (rc) return rc;
  uint32_t box_bytes = 2;
  for (int i = 0; i < 5; ++i) box_bytes *= box[i];
  const int smem = dump_bytes + 2048;
  CK(cudaFuncSetAttribute(probe_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_tma_kernel<<<1, 128, smem>>>(tm, coords[0], coords[1], coords[2], coords[3], coords[4], box_bytes, dump_bytes,
                                     d_out);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(out_host, d_out, dump_bytes, cudaMemcpyDeviceToHost));
  cudaFree(d_data);
  cudaFree(d_out);
  return 0;
}

// Returns achieved GB/s (smem fill rate summed over CTAs) or a negative error.
extern "C" double probe_tma_bw(uint64_t bytes, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle,
                               int n1, int n2, int n3, int s1, int s2, int s3, int ctas, int iters) {

  uint8_t* d_data;
  uint32_t* d_sink;
  if (cudaMalloc(&d_data, bytes) != cudaSuccess) return -1;
  cudaMemset(d_data, 1, bytes);
  cudaMalloc(&d_sink, ctas * 4);
  CUtensorMap tm;
  if (encode_map(&tm, d_data, 5, dims, strides_bytes, box, swizzle)) return -2;
  uint32_t box_bytes = 2;
  for (int i = 0; i < 5; ++i) box_bytes *= box[i];
  const int smem = 4 * ((box_bytes + 1023) & ~1023u) + 2048;
  if (cudaFuncSetAttribute(probe_tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
    return -3;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    probe_tma_bw_kernel<<<ctas, 128, smem>>>(tm, box_bytes, n1, n2, n3, s1, s2, s3, iters, d_sink);
    cudaEventRecord(e1);
    if (cudaDeviceSynchronize() != cudaSuccess) return -4;
  }
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaFree(d_data);
  cudaFree(d_sink);
  return (double)box_bytes * iters * ctas / (ms * 1e-3) / 1e9;
}
