// peer.cu -- the spike-tip exchange between neighbouring ranks through NVLink peer memory.
//
// The multi-GPU SPIKE path exchanges three small items per factor+solve across every rank boundary
// (include/spike_b200.h, SPK_BND_*): W^(t) of the right rank's first partition (kp*kp doubles, once per
// factorisation), g^(t) (kp doubles, per solve, right -> left) and x^(b) (kp doubles, per solve, left -> right).
// Through NCCL send/recv each of them costs a host round (get_boundary copy, ncclSend/Recv launch, stream events,
// set_boundary copy): ~80 us, three times per step, on a 2 ms step at 8 GPUs.  Here every rank owns a MAILBOX in its
// device memory, mapped into both neighbours with CUDA IPC; the producer's kernel stores the item straight into the
// consumer's mailbox over NVLink and releases a sequence flag, the consumer's kernel acquires the flag, copies
// the item where set_boundary would have put it and acknowledges into the producer's mailbox.  No host
// synchronisation, no second stream: post and wait are ordinary kernels on the engine's stream, so the W^(t)
// exchange still overlaps the band LU (posted before it, awaited after it).
//
// Mailbox layout (identical on all ranks; doubles, then 64-bit words):
//   channel 0 (W^(t),  right -> left): data[2][kp*kp]     channel 1 (g^(t), right -> left): data[2][kp*cols]
//   channel 2 (x^(b),  left -> right): data[2][kp*cols]   words: per channel flag[2], ack      (cols: spk_reserve_rhs, default 1)
// data/flag slots alternate with the sequence number; a producer may be at most two items ahead of its
// consumer (it spins on the ack word, which lives in ITS mailbox), so a rank that calls spk_factor repeatedly
// cannot overrun a slow neighbour.  Every spin is bounded (~2 s of SM clocks): a protocol error surfaces as
// SPK_ERR_STATE from spk_peer_check instead of a hung GPU.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

struct PeerLayout {
  size_t data_off[3];   // in doubles
  size_t count[3];
  size_t words_off;     // in doubles (= 8-byte words)
  size_t total;         // in doubles
};
static PeerLayout peer_layout(const spk_ctx* c) {
  PeerLayout L;
  const int kp = c->kp;
  const size_t kk = (size_t)kp * kp;
  L.count[0] = kk; L.count[1] = (size_t)kp * c->bnd_cols; L.count[2] = (size_t)kp * c->bnd_cols;
  size_t off = 0;
  for (int ch = 0; ch < 3; ++ch) { L.data_off[ch] = off; off += 2 * L.count[ch]; }
  L.words_off = off;
  L.total = off + 3 * 3 + 1;   // flag[2] + ack per channel, one error word
  return L;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Spin bound in SM clock ticks: SPIKE_B200_PEER_TIMEOUT_S seconds (default 20) at 2 GHz.  An expired spin is LOUD on
// the data path: the consumer fills its destination with NaN and does not acknowledge, the producer neither stores
// nor releases (its consumer then expires in turn); the sticky error word is mirrored into pinned host memory after
// every factor / solve (spk_peer_note) and fails the next call.
static long long peer_spin_cycles() {
  const char* e = getenv("SPIKE_B200_PEER_TIMEOUT_S");
  double sec = e ? atof(e) : 20.0;
  if (!(sec > 0.0)) sec = 20.0;
  return (long long)(sec * 2.0e9);
}

// producer: wait until the slot is free (ack >= seq-2), store the item into the consumer's mailbox, release the flag
__global__ void __launch_bounds__(256) k_peer_post(const double* __restrict__ src, int n, double* dst, unsigned long long* flag,
                                                   const unsigned long long* ack, unsigned long long seq, unsigned long long* err, long long limit) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const long long t0 = clock64();
    while (ld_acquire_sys(ack) + 2 < seq) {
      if (clock64() - t0 > limit) { atomicExch(err, 1ull); ok = 0; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!ok) return;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(flag, seq);
}

// consumer: acquire the flag, copy the item out of my mailbox, acknowledge into the producer's mailbox
__global__ void __launch_bounds__(256) k_peer_wait(double* __restrict__ dst, int n, const double* src, const unsigned long long* flag,
                                                   unsigned long long* ack, unsigned long long seq, unsigned long long* err, long long limit) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
      if (clock64() - t0 > limit) { atomicExch(err, 2ull); ok = 0; break; }
      __nanosleep(200);
    }
  }
  __syncthreads();
  if (!ok) {   // poison: whatever consumes this item must not look plausible
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __longlong_as_double(0x7ff8000000000000ll);
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldcv(src + i);
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(ack, seq);
}

extern "C" int spk_peer_mailbox_create(spk_ctx* c, void* handle64, void** dev_ptr) {
  if (!c || !c->have_band) return SPK_ERR_ARG;
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  const PeerLayout L = peer_layout(c);
  if (!c->mbox) {
    SPK_CUDA(c, cudaMalloc(&c->mbox, sizeof(double) * L.total));
    SPK_CUDA(c, cudaMemsetAsync(c->mbox, 0, sizeof(double) * L.total, c->stream));
    SPK_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int ch = 0; ch < 3; ++ch) c->peer_seq_out[ch] = c->peer_seq_in[ch] = 0;
    if (!c->h_peer_err) SPK_CUDA(c, cudaHostAlloc((void**)&c->h_peer_err, sizeof(unsigned long long), cudaHostAllocDefault));
    *c->h_peer_err = 0ull;   // a new mailbox starts clean
  }
  if (handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    SPK_CUDA(c, cudaIpcGetMemHandle(&h, c->mbox));
    memcpy(handle64, &h, 64);
  }
  if (dev_ptr) *dev_ptr = c->mbox;
  return SPK_OK;
}

// side 0: left neighbour, 1: right neighbour.  direct_ptr != NULL: the neighbour's mailbox lives in this process
// (R shards on one GPU, tests) and is used as it is; otherwise the IPC handle is opened.
extern "C" int spk_peer_mailbox_attach(spk_ctx* c, int side, const void* handle64, void* direct_ptr) {
  if (!c || side < 0 || side > 1 || (!handle64 && !direct_ptr)) return SPK_ERR_ARG;
  if (!c->mbox) { SPK_SET_ERR(c, "spk_peer_mailbox_attach: create the local mailbox first"); return SPK_ERR_STATE; }
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  if (direct_ptr) { c->peer_mbox[side] = (double*)direct_ptr; c->peer_ipc[side] = 0; return SPK_OK; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  SPK_CUDA(c, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_mbox[side] = (double*)p; c->peer_ipc[side] = 1;
  return SPK_OK;
}

void spk_peer_release(spk_ctx* c) {
  for (int s = 0; s < 2; ++s) {
    if (c->peer_mbox[s] && c->peer_ipc[s]) cudaIpcCloseMemHandle(c->peer_mbox[s]);
    c->peer_mbox[s] = nullptr; c->peer_ipc[s] = 0;
  }
  if (c->mbox) { cudaFree(c->mbox); c->mbox = nullptr; }
  if (c->h_peer_err) { cudaFreeHost(c->h_peer_err); c->h_peer_err = nullptr; }
}

// the error word as last mirrored into pinned host memory (no synchronisation): non-zero = an earlier exchange expired
int spk_peer_failed(spk_ctx* c) {
  if (!c->h_peer_err || !*(volatile unsigned long long*)c->h_peer_err) return 0;
  SPK_SET_ERR(c, "an earlier peer mailbox exchange timed out (code %llu): results since then are poisoned", *c->h_peer_err);
  return 1;
}
// enqueue the mirror copy (end of every factor / solve on a context with a mailbox)
int spk_peer_note(spk_ctx* c) {
  if (!c->mbox || !c->h_peer_err) return SPK_OK;
  const PeerLayout L = peer_layout(c);
  SPK_CUDA(c, cudaMemcpyAsync(c->h_peer_err, reinterpret_cast<unsigned long long*>(c->mbox + L.words_off) + 9, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  return SPK_OK;
}

static int peer_channel(int which, int* ch, int* side, int* is_post) {
  switch (which) {
    case SPK_BND_WT_FIRST:      *ch = 0; *side = 0; *is_post = 1; return 0;
    case SPK_BND_G_TOP:         *ch = 1; *side = 0; *is_post = 1; return 0;
    case SPK_BND_X_BOT:         *ch = 2; *side = 1; *is_post = 1; return 0;
    case SPK_BND_REMOTE_WT:     *ch = 0; *side = 1; *is_post = 0; return 0;
    case SPK_BND_REMOTE_G_TOP:  *ch = 1; *side = 1; *is_post = 0; return 0;
    case SPK_BND_REMOTE_X_BOT:  *ch = 2; *side = 0; *is_post = 0; return 0;
    default: return 1;
  }
}

// Send the boundary item `which` (an "out" item of SPK_BND_*) to the neighbour that consumes it.
extern "C" int spk_peer_post(spk_ctx* c, int which) {
  int ch, side, post;
  if (!c || peer_channel(which, &ch, &side, &post) || !post) return SPK_ERR_ARG;
  if (!c->mbox || !c->peer_mbox[side]) { SPK_SET_ERR(c, "spk_peer_post: no mailbox attached on side %d", side); return SPK_ERR_STATE; }
  if (spk_peer_failed(c)) return SPK_ERR_STATE;
  double* p; size_t n; int out;
  if (spk_bnd_desc(c, which, &p, &n, &out) || !out) { SPK_SET_ERR(c, "spk_peer_post: item %d unavailable", which); return SPK_ERR_ARG; }
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  if (which == SPK_BND_WT_FIRST) spk_side_join(c);   // W^(t) may have been computed on the side stream (capi.cu)
  const PeerLayout L = peer_layout(c);
  const unsigned long long seq = ++c->peer_seq_out[ch];
  const int slot = (int)(seq & 1ull);
  double* peer = c->peer_mbox[side];
  unsigned long long* peer_words = reinterpret_cast<unsigned long long*>(peer + L.words_off);
  unsigned long long* my_words = reinterpret_cast<unsigned long long*>(c->mbox + L.words_off);
  k_peer_post<<<1, 256, 0, c->stream>>>(p, (int)n, peer + L.data_off[ch] + (size_t)slot * L.count[ch], peer_words + 3 * ch + slot,
                                        my_words + 3 * ch + 2, seq, my_words + 9, peer_spin_cycles());
  SPK_KERNEL_CHECK(c);
  return SPK_OK;
}

// Receive the boundary item `which` (an "in" item of SPK_BND_*) from the neighbour that produces it.
extern "C" int spk_peer_wait(spk_ctx* c, int which) {
  int ch, side, post;
  if (!c || peer_channel(which, &ch, &side, &post) || post) return SPK_ERR_ARG;
  if (!c->mbox || !c->peer_mbox[side]) { SPK_SET_ERR(c, "spk_peer_wait: no mailbox attached on side %d", side); return SPK_ERR_STATE; }
  if (spk_peer_failed(c)) return SPK_ERR_STATE;
  double* p; size_t n; int out;
  if (spk_bnd_desc(c, which, &p, &n, &out) || out) { SPK_SET_ERR(c, "spk_peer_wait: bad item %d", which); return SPK_ERR_ARG; }
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  const PeerLayout L = peer_layout(c);
  const unsigned long long seq = ++c->peer_seq_in[ch];
  const int slot = (int)(seq & 1ull);
  unsigned long long* peer_words = reinterpret_cast<unsigned long long*>(c->peer_mbox[side] + L.words_off);
  unsigned long long* my_words = reinterpret_cast<unsigned long long*>(c->mbox + L.words_off);
  k_peer_wait<<<1, 256, 0, c->stream>>>(p, (int)n, c->mbox + L.data_off[ch] + (size_t)slot * L.count[ch], my_words + 3 * ch + slot,
                                        peer_words + 3 * ch + 2, seq, my_words + 9, peer_spin_cycles());
  SPK_KERNEL_CHECK(c);
  if (which == SPK_BND_REMOTE_WT) c->have_remote_wt = 1;
  return SPK_OK;
}

// Synchronises the engine's stream and reports a timed-out spin (1: a post waited for an ack, 2: a wait for a flag).
extern "C" int spk_peer_check(spk_ctx* c) {
  if (!c || !c->mbox) return SPK_ERR_ARG;
  SPK_CUDA(c, cudaSetDevice(c->opts.device));
  const PeerLayout L = peer_layout(c);
  unsigned long long e = 0;
  SPK_CUDA(c, cudaMemcpyAsync(&e, reinterpret_cast<unsigned long long*>(c->mbox + L.words_off) + 9, sizeof(e), cudaMemcpyDeviceToHost, c->stream));
  SPK_CUDA(c, cudaStreamSynchronize(c->stream));
  if (e) { SPK_SET_ERR(c, "peer mailbox exchange timed out (%s never arrived)", e == 1 ? "an acknowledgement" : "an item"); return SPK_ERR_STATE; }
  return SPK_OK;
}
