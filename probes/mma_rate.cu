// Probe: tcgen05.mma issue rate for (M=128, N, K=16) bf16 SS-mode instructions, as a function of N, the swizzle / row
// pitch of the K-major operands and the number of CTAs per SM.  Answers "what is the tensor-pipe ceiling of a C_out=32
// implicit-GEMM convolution" (hard part 1).  Diagnostic only; not part of the product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o probes/mma_rate probes/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../general-medical-image-segmentation-cnn-framework_b200/csrc/ptx.cuh"

using namespace b200;

struct Args {
  int N, KC, iters, nacc, a_tiles, a_sbo_rows, mn_major;
  long long* cycles;
};

__global__ void __launch_bounds__(128) mma_rate_kernel(Args p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int tid = threadIdx.x, warp = tid >> 5;
  const unsigned rowbytes = p.KC * 2;
  const unsigned a_bytes = 160 * p.a_sbo_rows / 8 * rowbytes;  // room for shifted views
  // fill operands with small non-zero values
  for (unsigned i = tid; i < (p.a_tiles * a_bytes + 256 * rowbytes) / 2; i += 128)
    reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(((i * 2654435761u) >> 28) * 0.125f - 1.f);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  if (warp == 1) {
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, p.N, p.mn_major, p.mn_major);
    const uint32_t swz = p.KC == 64 ? SWZ_128B : (p.KC == 32 ? SWZ_64B : SWZ_32B);
    // K-major: LBO ignored, SBO = 8-row group stride.  MN-major (weight-gradient form): rows are K (voxels), each row
    // holds KC channels of one M/N block; LBO = byte distance between blocks (here: one row, i.e. shifted views),
    // SBO = distance between 8-row K groups.
    const uint32_t a_hi = p.mn_major ? static_cast<uint32_t>(make_smem_desc(0, rowbytes, p.a_sbo_rows * rowbytes, swz) >> 32)
                                     : static_cast<uint32_t>(make_smem_desc(0, 16, p.a_sbo_rows * rowbytes, swz) >> 32);
    const uint32_t b_hi = p.mn_major ? static_cast<uint32_t>(make_smem_desc(0, 24 * rowbytes, 8 * rowbytes, swz) >> 32)
                                     : static_cast<uint32_t>(make_smem_desc(0, 16, 8 * rowbytes, swz) >> 32);
    const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + p.a_tiles * a_bytes);
    const int ksteps = p.KC / 16;
    const long long t0 = clock64();
    // descriptors are loop-invariant: the loop body is nothing but UTCHMMA issue slots (8 instructions per trip)
    const uint32_t a_lo = ((sA >> 4) & 0x3FFF) | (p.mn_major ? ((rowbytes >> 4) << 16) : (1u << 16));
    const uint32_t b_lo = ((sB >> 4) & 0x3FFF) | (p.mn_major ? (((24 * rowbytes) >> 4) << 16) : (1u << 16));
    const uint64_t ad0 = (static_cast<uint64_t>(a_hi) << 32) | a_lo, ad1 = ad0 + 2 + (p.a_tiles > 1 ? 8 * p.a_sbo_rows * rowbytes / 16 : 0);
    const uint64_t bd0 = (static_cast<uint64_t>(b_hi) << 32) | b_lo, bd1 = bd0 + 2;
    (void)ksteps;
    for (int it = 0; it < p.iters; it += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u)
        umma_f16_pred(tbase + (u % p.nacc) * p.N, (u & 1) ? ad1 : ad0, (u & 1) ? bd1 : bd0, idesc, 1u, leader);
    }
    umma_commit_pred(&bar, leader);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if ((tid & 31) == 0) p.cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tbase, 512);
  }
}

int main() {
  long long* d_cycles;
  cudaMalloc(&d_cycles, 1024 * sizeof(long long));
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4096;
  printf("%-6s %-4s %-4s %-7s %-5s %-6s | cyc/MMA  TFLOP/s(chip @ measured clock)\n", "N", "KC", "nacc", "a_tiles", "sbo", "ctas");
  struct Case { int N, KC, nacc, a_tiles, sbo, ctas_per_sm; };
  std::vector<Case> cases;
  for (int N : {16, 32, 64, 128, 256})
    for (int KC : {32, 64}) {
      cases.push_back({N, KC, 512 / N > 8 ? 8 : 512 / N, 1, 8, 1});
      cases.push_back({N, KC, 512 / N > 8 ? 8 : 512 / N, 4, 10, 1});
    }
  // MN-major (wgrad form), marked by ctas_per_sm = -1: N = blocks x KC channels
  for (int KC : {32, 64})
    for (int N : {KC, 2 * KC, 3 * KC, 4 * KC}) {
      if (N > 256) continue;
      cases.push_back({N, KC, 512 / N > 4 ? 4 : 512 / N, 1, 10, -1});
    }
  for (const Case& c : cases) {
    Args a{c.N, c.KC, iters, c.nacc, c.a_tiles, c.sbo, c.ctas_per_sm < 0 ? 1 : 0, d_cycles};
    const unsigned rowbytes = c.KC * 2;
    const size_t smem = static_cast<size_t>(c.a_tiles) * (160 * c.sbo / 8 * rowbytes) + 256 * rowbytes + 2048;
    const int ctas = 148;
    // with 2 CTAs per SM each CTA can only own 256 TMEM columns; the kernel allocates 512, so emulate by halving
    if (c.ctas_per_sm > 1) continue;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    mma_rate_kernel<<<ctas, 128, smem>>>(a);
    cudaEventRecord(e0);
    mma_rate_kernel<<<ctas, 128, smem>>>(a);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("N %d KC %d: %s\n", c.N, c.KC, cudaGetErrorString(err));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), d_cycles, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long v : h) avg += v;
    avg /= ctas;
    const double mmas = static_cast<double>(iters);
    const double flops = mmas * 2.0 * 128 * c.N * 16 * ctas;
    printf("%s%-6d %-4d %-4d %-7d %-5d %-6d | %7.1f  %8.1f  (kernel %.3f ms)\n", c.ctas_per_sm < 0 ? "MN " : "", c.N, c.KC, c.nacc, c.a_tiles, c.sbo, ctas,
           avg / mmas, flops / (ms * 1e-3) / 1e12, ms);
  }
  return 0;
}
