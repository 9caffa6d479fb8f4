// Small all-reduce over NVLink peer memory, fused with the batch-norm statistics finalisation.
//
// SynchronizedBatchNorm needs, per layer and per step, the sum over ranks of 2*C+1 floats (sum, sum of squares, element
// count) in the forward pass and of 2*C floats in the backward pass: 36 latency-bound exchanges per U-Net step (18 forward, 18 backward), on the
// critical path (the reference does them with Python threads, queues and two coalesced device copies per layer:
// models/sync_batchnorm/batchnorm.py:90-111, comm.py:56-137).  An NCCL call per exchange costs a launch plus a
// protocol round trip, and NCCL collectives cannot be captured into the training step's CUDA graph on this stack.
// Here every rank owns a mailbox in its own HBM that its peers map through CUDA IPC.  One single-CTA kernel
//   1. stores its vector into slot [seq % SLOTS][rank] of EVERY peer's mailbox over NVLink, every float as one 8-byte
//      {value, seq} word (no separate flag and no fence: the receiver polls the words themselves),
//   2. spins (bounded) until its own mailbox shows seq-tagged words from every rank, sums them in rank order -- the
//      same order everywhere, so all ranks obtain bit-identical results -- and
//   3. optionally finalises the statistics in the same launch (mean, inv_std, fused scale / shift, running stats).
// `seq` lives in device memory and is advanced by the kernel, so a captured launch replays correctly.  A rank can be at
// most one exchange ahead of the slowest one (it cannot finish exchange k+1 before everybody has written it, which they
// do only after finishing k), hence SLOTS >= 2 makes slot reuse safe; 4 are used.
#include <string.h>

#include "common.cuh"

namespace b200 {

constexpr int kP2PSlots = 4;
constexpr int kP2PMaxWorld = 8;
constexpr int kP2PMaxFloats = 2112;   // 2 * 1024 channels + count, padded

// Every float travels as one 8-byte word {value bits, sequence number}: the receiver polls the word itself, so no
// separate flag, no system fence between data and flag and no second NVLink transaction are needed (the idea of NCCL's LL
// protocol; an aligned 8-byte store is a single transaction).  A slot written in exchange `seq` is reused in exchange
// seq + kP2PSlots, whose tag differs, so stale words are never mistaken for fresh ones.
struct P2PMailbox {
  uint2 data[kP2PSlots][kP2PMaxWorld][kP2PMaxFloats];
};

struct P2PPeers {
  P2PMailbox* box[kP2PMaxWorld];
};

__device__ __forceinline__ void st_ll(uint2* p, float v, unsigned int tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint2 ld_ll(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// vec: n floats, reduced in place.  finalize != 0: vec = {sum[C], sumsq[C], count} and coef / running stats are produced
// exactly like norm_finalize_kernel does from the reduced sums.
// phase: 0 = the whole exchange in one launch; 1 = steps 1-2 only (send: returns as soon as the vector is on its way);
// 2 = steps 3-4 only (receive).  Launching 1, then unrelated kernels, then 2 hides the NVLink round trip behind them.
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(float* __restrict__ vec, int n, P2PPeers peers, int rank,
                                                            int world, unsigned int* __restrict__ seq_ptr, int phase, int finalize,
                                                            float local_count, int C, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ running_mean,
                                                            float* __restrict__ running_var, float momentum, float eps,
                                                            int clamp_eps, float* __restrict__ coef) {
  const unsigned int seq = *seq_ptr + 1;
  const int slot = seq % kP2PSlots;
  if (phase != 2) {
    if (finalize) {   // the element count travels with the sums (sum_size in batchnorm.py:58-62)
      if (threadIdx.x == 0) vec[2 * C] = local_count;
      __syncthreads();
    }
    // 1 + 2. scatter my vector, tagged word by word, into every rank's mailbox (my own included)
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = vec[i];
      for (int r = 0; r < world; ++r) st_ll(&peers.box[r]->data[slot][rank][i], v, seq);
    }
    if (phase == 1) return;
    __syncthreads();     // (phase 0: vec is overwritten below by other threads' sums only at their own indices: no hazard,
                         //  but keep the block together so that the spin starts after all stores were issued)
  }
  // 3. wait for everybody's words and sum them in rank order -- the same order everywhere: bit-identical results
  P2PMailbox* mine = peers.box[rank];
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) {
      uint2 w = ld_ll(&mine->data[slot][r][i]);
      while (w.y != seq) {
        if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s: a peer never arrived (mismatched call sequence)
        w = ld_ll(&mine->data[slot][r][i]);
      }
      s += __uint_as_float(w.x);
    }
    vec[i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) *seq_ptr = seq;
  if (!finalize) return;
  // 4. statistics -> normalisation constants (same arithmetic as norm_finalize_kernel, elementwise.cu)
  const double count = vec[2 * C];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double s = vec[c], q = vec[C + c];
    const double mean = s / count;
    double var = (q - s * mean) / count;  // batchnorm.py:116-120: sumvar = ssum - sum*mean
    if (var < 0) var = 0;
    const double inv_std = clamp_eps ? 1.0 / sqrt(var < eps ? static_cast<double>(eps) : var) : 1.0 / sqrt(var + eps);
    if (running_mean != nullptr) {
      const double unbiased = count > 1 ? var * count / (count - 1) : var;
      running_mean[c] = static_cast<float>((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = static_cast<float>((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
    const double ga = gamma ? gamma[c] : 1.0;
    const double be = beta ? beta[c] : 0.0;
    coef[0 * C + c] = static_cast<float>(mean);
    coef[1 * C + c] = static_cast<float>(inv_std);
    coef[2 * C + c] = static_cast<float>(ga * inv_std);
    coef[3 * C + c] = static_cast<float>(be - mean * ga * inv_std);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Gradient all-reduce over NVLink peer memory, capturable in the training step's CUDA graph (train.py:211: the all-reduce
// DDP performs inside accelerator.backward).  Every rank's gradient buffer (optim.FusedAdam's zeroed arena: torch-layout
// gradients followed by the conv kernels' packed weight-gradient accumulators), a staging buffer and a small flag block
// live in cudaMalloc'ed memory that all peers map through CUDA IPC.  Only the 64-float chunks that ever receive a gradient
// are exchanged (`live`: sorted chunk indices, identical on every rank).  Four launches, all stream-ordered:
//   1. signal A: "my gradients of exchange `seq` are complete" -> every peer's flag block (release, system scope);
//   2. reduce + push: wait for A from everybody; rank r sums ITS share of the live chunks over all ranks' buffers in rank
//      order (peer loads over NVLink; the same order everywhere) and stores the sum into EVERY rank's buffer;
//   3. signal B: "my share has been pushed" (system-scope release after the pushes);
//   4. wait for B from everybody.  All ranks end with bit-identical sums.
// Hazards: during an exchange the chunks of rank r's share are read and written by rank r only, in all buffers; a buffer is
// next overwritten by its owner's following backward pass, which starts after its launch 4, i.e. after every rank has
// finished reading and pushing.  `seq` lives in device memory and is advanced by launch 1, so a captured sequence replays
// correctly.  (`reds`, a staging buffer, is no longer used by the kernels; the ABI keeps the argument.)
constexpr int kGradChunk = 64;   // floats

struct GradPeers {
  float* buf[kP2PMaxWorld];
  float* red[kP2PMaxWorld];
  unsigned int* flags[kP2PMaxWorld];   // [2][kP2PMaxWorld] per rank
};

__global__ void p2p_grad_signal_kernel(GradPeers peers, int which, int rank, int world, unsigned int* __restrict__ seq_ptr,
                                       int bump) {
  __shared__ unsigned int seq;
  if (threadIdx.x == 0) {
    seq = *seq_ptr + (bump ? 1u : 0u);
    if (bump) *seq_ptr = seq;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) st_release_sys(&peers.flags[threadIdx.x][which * kP2PMaxWorld + rank], seq);
}

__device__ __forceinline__ void grad_wait(const unsigned int* mine, int which, int world, unsigned int seq) {
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(&mine[which * kP2PMaxWorld + threadIdx.x]) < seq) {
      if (clock64() - t0 > 8000000000LL) __trap();   // ~4 s: a peer never arrived
    }
  }
  __syncthreads();
}

// Reduce + broadcast: rank r sums its share of the live chunks over all ranks' buffers (W peer loads in flight per thread,
// rank order) and PUSHES the sum into every rank's buffer.  Nobody else touches those chunks during the exchange (each
// rank reads / writes only its own share, in everybody's buffer), so the pushes need no staging copy; stores over NVLink
// are fire-and-forget, the all-gather half costs no round trip.
template <int W>
__global__ void __launch_bounds__(256) p2p_grad_reduce_push_kernel(GradPeers peers, const int* __restrict__ live, int n_live,
                                                                   int rank, const unsigned int* __restrict__ seq_ptr) {
  grad_wait(peers.flags[rank], 0, W, *seq_ptr);
  const int per = (n_live + W - 1) / W;
  const int lo = rank * per, hi = min(n_live, lo + per);
  const long long items = static_cast<long long>(max(hi - lo, 0)) * (kGradChunk / 4);
  for (long long w = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; w < items;
       w += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int li = lo + static_cast<int>(w / (kGradChunk / 4)), q = static_cast<int>(w % (kGradChunk / 4));
    const long long off = static_cast<long long>(live[li]) * kGradChunk + q * 4;
    float4 v[W];
#pragma unroll
    for (int r = 0; r < W; ++r) v[r] = __ldcv(reinterpret_cast<const float4*>(peers.buf[r] + off));
    float4 acc = v[0];
#pragma unroll
    for (int r = 1; r < W; ++r) {
      acc.x += v[r].x;
      acc.y += v[r].y;
      acc.z += v[r].z;
      acc.w += v[r].w;
    }
#pragma unroll
    for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(peers.buf[r] + off) = acc;
  }
  __threadfence_system();     // the pushes are performed system-wide before this thread retires (flag B follows the kernel)
}

// Completion: everybody's pushes have landed (flag B from every rank, sent after its reduce+push kernel drained).
__global__ void p2p_grad_wait_kernel(GradPeers peers, int rank, int world, const unsigned int* __restrict__ seq_ptr) {
  grad_wait(peers.flags[rank], 1, world, *seq_ptr);
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200seg_p2p_mailbox_bytes(void) { return sizeof(P2PMailbox); }

int b200seg_p2p_alloc(void** dev_ptr, void* ipc_handle_out) {
  B200_CHECK_ARG(dev_ptr && ipc_handle_out, "p2p_alloc: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, sizeof(P2PMailbox));
  if (e == cudaSuccess) e = cudaMemset(p, 0, sizeof(P2PMailbox));
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("p2p_alloc: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    return B200SEG_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

int b200seg_p2p_alloc_bytes(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
  B200_CHECK_ARG(dev_ptr && ipc_handle_out && bytes > 0, "p2p_alloc_bytes: bad arguments");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("p2p_alloc_bytes: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    return B200SEG_ERR_CUDA;
  }
  memcpy(ipc_handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return 0;
}

int b200seg_p2p_grad_allreduce(const void* const* bufs, const void* const* reds, const void* const* flags, const int32_t* live,
                               int n_live, int rank, int world, uint32_t* seq, void* stream) {
  B200_CHECK_ARG(bufs && reds && flags && live && seq && n_live > 0, "p2p_grad_allreduce: bad arguments");
  B200_CHECK_ARG(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "p2p_grad_allreduce: bad rank / world");
  GradPeers peers{};
  for (int r = 0; r < world; ++r) {
    B200_CHECK_ARG(bufs[r] && reds[r] && flags[r], "p2p_grad_allreduce: null peer pointer for rank %d", r);
    peers.buf[r] = static_cast<float*>(const_cast<void*>(bufs[r]));
    peers.red[r] = static_cast<float*>(const_cast<void*>(reds[r]));
    peers.flags[r] = static_cast<unsigned int*>(const_cast<void*>(flags[r]));
  }
  auto st = static_cast<cudaStream_t>(stream);
  const int per = (n_live + world - 1) / world;
  const int g_rs = grid_for(static_cast<int64_t>(per) * (kGradChunk / 4), 256, kNumSMs * 4);
  p2p_grad_signal_kernel<<<1, 32, 0, st>>>(peers, 0, rank, world, seq, 1);
  switch (world) {
    case 1: p2p_grad_reduce_push_kernel<1><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 2: p2p_grad_reduce_push_kernel<2><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 3: p2p_grad_reduce_push_kernel<3><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 4: p2p_grad_reduce_push_kernel<4><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 5: p2p_grad_reduce_push_kernel<5><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 6: p2p_grad_reduce_push_kernel<6><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    case 7: p2p_grad_reduce_push_kernel<7><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
    default: p2p_grad_reduce_push_kernel<8><<<g_rs, 256, 0, st>>>(peers, live, n_live, rank, seq); break;
  }
  p2p_grad_signal_kernel<<<1, 32, 0, st>>>(peers, 1, rank, world, seq, 0);
  p2p_grad_wait_kernel<<<1, 32, 0, st>>>(peers, rank, world, seq);
  B200_CHECK_LAUNCH("p2p_grad_allreduce");
  return 0;
}

int b200seg_p2p_open(const void* ipc_handle, void** dev_ptr) {
  B200_CHECK_ARG(ipc_handle && dev_ptr, "p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof(h));
  const cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("p2p_open: %s", cudaGetErrorString(e));
    return B200SEG_ERR_CUDA;
  }
  return 0;
}

int b200seg_p2p_close(void* dev_ptr, int opened) {
  const cudaError_t e = opened ? cudaIpcCloseMemHandle(dev_ptr) : cudaFree(dev_ptr);
  if (e != cudaSuccess) {
    set_error("p2p_close: %s", cudaGetErrorString(e));
    return B200SEG_ERR_CUDA;
  }
  return 0;
}

int b200seg_p2p_allreduce(float* vec, int n, const void* const* mailboxes, int rank, int world, uint32_t* seq, int phase,
                          int finalize_channels, double local_count, const float* gamma, const float* beta, float* running_mean,
                          float* running_var, float momentum, float eps, int clamp_eps, float* coef, void* stream) {
  B200_CHECK_ARG(vec && mailboxes && seq && n > 0 && n <= kP2PMaxFloats, "p2p_allreduce: at most %d floats", kP2PMaxFloats);
  B200_CHECK_ARG(world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "p2p_allreduce: bad rank / world");
  B200_CHECK_ARG(phase >= 0 && phase <= 2, "p2p_allreduce: phase must be 0 (whole exchange), 1 (send) or 2 (receive)");
  B200_CHECK_ARG(finalize_channels == 0 || (coef && n == 2 * finalize_channels + 1), "p2p_allreduce: finalize needs "
                 "{sum[C], sumsq[C], count} and a coefficient buffer");
  P2PPeers peers{};
  for (int r = 0; r < world; ++r) {
    B200_CHECK_ARG(mailboxes[r] != nullptr, "p2p_allreduce: null mailbox for rank %d", r);
    peers.box[r] = static_cast<P2PMailbox*>(const_cast<void*>(mailboxes[r]));
  }
  p2p_allreduce_kernel<<<1, 512, 0, static_cast<cudaStream_t>(stream)>>>(
      vec, n, peers, rank, world, seq, phase, finalize_channels > 0, static_cast<float>(local_count), finalize_channels, gamma, beta, running_mean, running_var,
      momentum, eps, clamp_eps, coef);
  B200_CHECK_LAUNCH("p2p_allreduce");
  return 0;
}

}  // extern "C"
