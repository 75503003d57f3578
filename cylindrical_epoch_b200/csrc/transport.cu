// transport.cu -- the x-neighbour exchange (MPI_SENDRECV in the reference: boundary.F90:528,
// 541,1187,1195; partlist.F90:842,869) for one handle.
//   NONE     : nranks == 1; a periodic domain wraps onto itself with device copies
//   NCCL     : ncclSend / ncclRecv inside one group on the library stream (NVLink 5);
//              libnccl is dlopen()ed so the library has no link-time NCCL dependency and
//              shares the NCCL already loaded by the host process (e.g. torch's)
//   CALLBACK : caller-supplied sendrecv (torch.distributed, MPI, ...)
//   FABRIC   : several handles in ONE process, one host thread each (single-GPU tests of
//              the multi-rank logic)
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include <condition_variable>
#include <cstring>
#include <mutex>

#include "ctx.cuh"

namespace cylgpu {

// ---- peer-memory mailboxes over NVLink (the data path of the NCCL transport) ----
// Every rank owns one device allocation with two mailboxes -- messages arriving from the left and from the right
// neighbour -- of two slots each, and maps its neighbours' allocations through CUDA IPC (handles travel once over
// the NCCL communicator).  A message is ONE kernel on the sender (copy into the neighbour's slot over NVLink,
// system-scope fence, flag = sequence number) and ONE on the receiver (wait for the flag, copy out, acknowledge):
// no rendezvous, no proxy thread, ~3 us instead of ~20 us per exchange, and capturable in CUDA graphs because the
// sequence numbers live in device memory.  A slab of an 8-GPU run makes ~10 exchanges per 3.5 ms step.
#define P2P_FLAGS_BYTES 256
// A wait gives up after ~2 minutes (a sticky error instead of a hung device).  Generous on purpose: the ranks of a run
// reach their first exchange as far apart as their set-up times differ (seconds), and NCCL would simply wait.
#define P2P_TIMEOUT_CLOCKS 240000000000LL   // ~120 s at 2 GHz
struct P2PBoxHeader {               // at the start of each mailbox (written by the PEER unless noted)
  unsigned long long flag[2];       // sequence number of the message in slot 0 / 1
  unsigned long long ack;           // highest sequence number the OWNER of the box the peer writes to has consumed:
                                    // written by the peer into MY box header about messages I sent
  unsigned long long pad[5];
};
struct P2P {
  bool ready = false;
  size_t cap = 0;                   // bytes per slot
  unsigned char* mine = nullptr;    // [from_left box][from_right box]; box = header (256 B) + 2 * cap
  unsigned char* peer_l = nullptr;  // the left neighbour's allocation (mapped): I write into its from_right box
  unsigned char* peer_r = nullptr;  // the right neighbour's allocation: I write into its from_left box
  bool same_peer = false;           // two ranks in a ring: left == right, one mapping
  unsigned long long* seq = nullptr;   // device: [0] sent left, [1] sent right, [2] received from left, [3] from right
  unsigned int* done = nullptr;        // device: block counters of the send kernels [2]
  int* status = nullptr;               // device: != 0 after a receive timed out
};

struct Transport {
  int kind = CYLGPU_TRANSPORT_NONE;
  P2P p2p;
  // NCCL
  void* nccl_lib = nullptr;
  void* comm = nullptr;
  void* comm2 = nullptr;   // a second communicator (ncclCommSplit) for exchanges enqueued on the side stream
  int (*ncclSend)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ncclRecv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ncclGroupStart)() = nullptr;
  int (*ncclGroupEnd)() = nullptr;
  int (*ncclCommDestroy)(void*) = nullptr;
  const char* (*ncclGetErrorString)(int) = nullptr;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// One leg of an exchange through the mailboxes.  send: src -> the peer's box, hdr = the header of MY box for messages
// from that neighbour (it holds the neighbour's acknowledgements); receive: my box -> dst, hdr = the PEER's header
// that takes my acknowledgement.  seq_ctr: my count of messages on this leg (device memory, so that a captured graph
// can be replayed).
struct P2PLeg {
  const double* src; double* dst; size_t n8;
  unsigned char* box; unsigned long long* seq_ctr; P2PBoxHeader* hdr; unsigned int* done;
};

// words of a message that carry data: all of it, or -- a particle message, [7-double header: count][slots] -- the
// header and the slots in use
__device__ __forceinline__ size_t p2p_live_words(double count_word, size_t n8, int counted) {
  if (!counted) return n8;
  const long long bits = __double_as_longlong(count_word);   // (the count travels as an int64 in the first word)
  const size_t want = 7 + 7 * (size_t)(bits > 0 ? bits : 0);
  return want < n8 ? want : n8;
}

// Both sends of an exchange in one launch (blockIdx.y = leg): wait until the slot is free (the receiver acknowledged
// the message two back), copy, publish.
__global__ void __launch_bounds__(256) k_p2p_send(P2PLeg legs0, P2PLeg legs1, size_t cap, int counted, int* status) {
  const P2PLeg L = blockIdx.y == 0 ? legs0 : legs1;
  if (L.n8 == 0) return;
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) {
    const unsigned long long seq = ld_volatile_u64(L.seq_ctr) + 1ULL;
    const long long t0 = clock64();
    while (seq > 2 && ld_volatile_u64(&L.hdr->ack) + 2ULL < seq) {
      if (clock64() - t0 > P2P_TIMEOUT_CLOCKS) { *status = 1; break; }   // never hang the device: report instead
    }
    s_seq = seq;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const size_t n = p2p_live_words(L.src[0], L.n8, counted);
  double* dst = reinterpret_cast<double*>(L.box + P2P_FLAGS_BYTES + (seq & 1ULL) * cap);
  if (((reinterpret_cast<uintptr_t>(L.src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(L.src);
    double2* d2 = reinterpret_cast<double2*>(dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (size_t)gridDim.x * blockDim.x)
      d2[i] = s2[i];
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) dst[n - 1] = L.src[n - 1];
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = L.src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(L.done, 1u);
    if (t == gridDim.x - 1) {      // the last block: everything is on its way, publish
      *L.done = 0u;
      *L.seq_ctr = seq;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(&reinterpret_cast<P2PBoxHeader*>(L.box)->flag[seq & 1ULL]) = seq;
    }
  }
}

// ... and both arrivals: wait for the flag of the next sequence number in MY mailbox, copy out (L2 loads: the data
// came from the peer), acknowledge in the PEER's header.
__global__ void __launch_bounds__(256) k_p2p_recv(P2PLeg legs0, P2PLeg legs1, size_t cap, int counted, int* status) {
  const P2PLeg L = blockIdx.y == 0 ? legs0 : legs1;
  if (L.n8 == 0) return;
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) {
    const unsigned long long seq = ld_volatile_u64(L.seq_ctr) + 1ULL;
    const P2PBoxHeader* hdr = reinterpret_cast<const P2PBoxHeader*>(L.box);
    const long long t0 = clock64();
    while (ld_volatile_u64(&hdr->flag[seq & 1ULL]) != seq) {
      if (clock64() - t0 > P2P_TIMEOUT_CLOCKS) { *status = 2; break; }
    }
    s_seq = seq;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const double* src = reinterpret_cast<const double*>(L.box + P2P_FLAGS_BYTES + (seq & 1ULL) * cap);
  const size_t n = p2p_live_words(__ldcg(src), L.n8, counted);
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(L.dst)) & 15) == 0) {
    const double2* s2 = reinterpret_cast<const double2*>(src);
    double2* d2 = reinterpret_cast<double2*>(L.dst);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2; i += (size_t)gridDim.x * blockDim.x)
      d2[i] = __ldcg(s2 + i);
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) L.dst[n - 1] = __ldcg(src + n - 1);
  } else {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      L.dst[i] = __ldcg(src + i);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(L.done, 1u);
    if (t == gridDim.x - 1) {
      *L.done = 0u;
      *L.seq_ctr = seq;
      __threadfence_system();
      *reinterpret_cast<volatile unsigned long long*>(&L.hdr->ack) = seq;
    }
  }
}

static void p2p_release(P2P& P) {
  if (P.peer_l) cudaIpcCloseMemHandle(P.peer_l);
  if (P.peer_r && !P.same_peer) cudaIpcCloseMemHandle(P.peer_r);
  P.peer_l = P.peer_r = nullptr;
  if (P.mine) cudaFree(P.mine);
  if (P.seq) cudaFree(P.seq);
  if (P.done) cudaFree(P.done);
  if (P.status) cudaFree(P.status);
  P.mine = nullptr; P.seq = nullptr; P.done = nullptr; P.status = nullptr;
  P.ready = false;
}

// (Re)build the mailboxes for messages of up to `cap` bytes.  Collective over the neighbours: every rank calls it
// at the same point (cylgpu_create, cylgpu_set_exchange_capacity).  Any failure leaves the NCCL path in charge.
int p2p_setup(cylgpu_ctx* c, size_t cap) {
  Transport* t = c->tr;
  c->p2p_link_l = c->p2p_link_r = false;
  if (!t || t->kind != CYLGPU_TRANSPORT_NCCL) return 0;
  // On by default for every message that fits a slot; CYLGPU_P2P=0: ncclSend / ncclRecv only, CYLGPU_P2P=particles:
  // the counted particle messages only.  Measured on C3 over 8 B200s with the communication-avoiding field phases
  // (profiles/r2g_*): 3.46 ms per step through the mailboxes (3.47 with the particle messages only) against 3.98 ms
  // through NCCL alone.  The first version (round 2, r2d: four launches per exchange, 32 blocks, whole fixed-size
  // particle messages) was slower than NCCL (3.81 against 3.73 ms); what changed: both legs of an exchange in one
  // launch, 16-byte copies on up to 64 blocks, and only the slots in use of a particle message cross the link.
  {
    const char* e = getenv("CYLGPU_P2P");
    c->p2p_policy = !e ? 1 : (strcmp(e, "particles") == 0 ? 2 : (atoi(e) != 0 ? 1 : 0));
    if (c->p2p_policy == 0) return 0;
  }
  const int left = c->left, right = c->right, me = c->cfg.rank;
  if ((left < 0 && right < 0) || left == me || right == me) return 0;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  P2P& P = t->p2p;
  p2p_release(P);
  cap = (cap + 255) / 256 * 256;
  const size_t box = P2P_FLAGS_BYTES + 2 * cap;
  bool ok = true;
  ok = ok && cudaMalloc(&P.mine, 2 * box) == cudaSuccess;
  ok = ok && cudaMemset(P.mine, 0, 2 * box) == cudaSuccess;
  ok = ok && cudaMalloc(&P.seq, 4 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMemset(P.seq, 0, 4 * sizeof(unsigned long long)) == cudaSuccess;
  ok = ok && cudaMalloc(&P.done, 4 * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMemset(P.done, 0, 4 * sizeof(unsigned int)) == cudaSuccess;
  ok = ok && cudaMalloc(&P.status, sizeof(int)) == cudaSuccess;
  ok = ok && cudaMemset(P.status, 0, sizeof(int)) == cudaSuccess;
  cudaIpcMemHandle_t mine_h;
  memset(&mine_h, 0, sizeof(mine_h));
  ok = ok && cudaIpcGetMemHandle(&mine_h, P.mine) == cudaSuccess;
  // handles (+ a validity byte) to both neighbours over NCCL
  struct Wire { cudaIpcMemHandle_t h; unsigned long long ok; unsigned long long cap; };
  Wire w_me;
  w_me.h = mine_h; w_me.ok = ok ? 1ULL : 0ULL; w_me.cap = cap;
  Wire* d_w = nullptr;   // [0] mine, [1] from left, [2] from right
  if (cudaMalloc(&d_w, 3 * sizeof(Wire)) != cudaSuccess) { cudaGetLastError(); p2p_release(P); return 0; }
  cudaMemset(d_w, 0, 3 * sizeof(Wire));
  cudaMemcpy(d_w, &w_me, sizeof(Wire), cudaMemcpyHostToDevice);
  int r = t->ncclGroupStart();
  if (r == 0 && left >= 0) r = t->ncclSend(d_w, sizeof(Wire), 1, left, t->comm, c->stream);
  if (r == 0 && right >= 0) r = t->ncclSend(d_w, sizeof(Wire), 1, right, t->comm, c->stream);
  if (r == 0 && right >= 0) r = t->ncclRecv(d_w + 2, sizeof(Wire), 1, right, t->comm, c->stream);
  if (r == 0 && left >= 0) r = t->ncclRecv(d_w + 1, sizeof(Wire), 1, left, t->comm, c->stream);
  const int r2 = t->ncclGroupEnd();
  if (r == 0) r = r2;
  Wire w_in[3];
  memset(w_in, 0, sizeof(w_in));
  if (r == 0 && cudaStreamSynchronize(c->stream) == cudaSuccess)
    cudaMemcpy(w_in, d_w, 3 * sizeof(Wire), cudaMemcpyDeviceToHost);
  cudaFree(d_w);
  if (r != 0) { set_error("NCCL exchange of the IPC handles failed"); p2p_release(P); return 5; }
  // a link is usable iff BOTH of its ends exported, got a valid handle of the same slot size and mapped it; each
  // link is judged on its own (the two links of a rank may differ: the exchange then mixes mailboxes and NCCL)
  P.same_peer = (left >= 0 && left == right);
  bool ok_l = ok && left >= 0 && w_in[1].ok == 1ULL && w_in[1].cap == cap;
  bool ok_r = ok && right >= 0 && w_in[2].ok == 1ULL && w_in[2].cap == cap;
  if (ok_l) ok_l = cudaIpcOpenMemHandle((void**)&P.peer_l, w_in[1].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  if (!ok_l) P.peer_l = nullptr;
  if (ok_r) {
    if (P.same_peer) { P.peer_r = P.peer_l; ok_r = ok_l; }
    else ok_r = cudaIpcOpenMemHandle((void**)&P.peer_r, w_in[2].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
  }
  if (!ok_r) P.peer_r = nullptr;
  cudaGetLastError();
  // second round: my verdict on each link to the neighbour at its other end
  unsigned long long* d_f = nullptr;   // [0] my verdict left link, [1] my verdict right link, [2] from left, [3] from right
  unsigned long long f_host[4] = {ok_l ? 1ULL : 0ULL, ok_r ? 1ULL : 0ULL, 0ULL, 0ULL};
  if (cudaMalloc(&d_f, sizeof(f_host)) != cudaSuccess) { cudaGetLastError(); p2p_release(P); return 0; }
  cudaMemcpy(d_f, f_host, sizeof(f_host), cudaMemcpyHostToDevice);
  r = t->ncclGroupStart();
  if (r == 0 && left >= 0) r = t->ncclSend(d_f + 0, 8, 1, left, t->comm, c->stream);
  if (r == 0 && right >= 0) r = t->ncclSend(d_f + 1, 8, 1, right, t->comm, c->stream);
  if (r == 0 && right >= 0) r = t->ncclRecv(d_f + 3, 8, 1, right, t->comm, c->stream);
  if (r == 0 && left >= 0) r = t->ncclRecv(d_f + 2, 8, 1, left, t->comm, c->stream);
  const int r3 = t->ncclGroupEnd();
  if (r == 0) r = r3;
  if (r == 0 && cudaStreamSynchronize(c->stream) == cudaSuccess)
    cudaMemcpy(f_host, d_f, sizeof(f_host), cudaMemcpyDeviceToHost);
  cudaFree(d_f);
  if (r != 0) { set_error("NCCL exchange of the mailbox verdicts failed"); p2p_release(P); return 5; }
  c->p2p_link_l = ok_l && f_host[2] == 1ULL;
  c->p2p_link_r = ok_r && f_host[3] == 1ULL;
  P.ready = c->p2p_link_l || c->p2p_link_r;
  P.cap = cap;
  return 0;
}

// ---- in-process fabric ----
struct Fabric {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  struct Slot {
    const void* send_l; size_t send_l_b;
    const void* send_r; size_t send_r_b;
  };
  std::vector<Slot> slots;
  explicit Fabric(int n_) : n(n_), slots(n_) {}
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long gen = generation;
    if (++arrived == n) {
      arrived = 0;
      ++generation;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != gen; });
    }
  }
};

static void* open_nccl() {
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  return h;
}

struct NcclUniqueId { char internal[128]; };

Transport* make_transport(cylgpu_ctx* c) {
  Transport* t = new Transport();
  t->kind = c->cfg.transport;
  if (c->cfg.nranks > 1 && t->kind == CYLGPU_TRANSPORT_NONE) {
    set_error("nranks > 1 needs a transport (NCCL, CALLBACK or FABRIC)");
    delete t;
    return nullptr;
  }
  if (t->kind == CYLGPU_TRANSPORT_NCCL) {
    t->nccl_lib = open_nccl();
    if (!t->nccl_lib) {
      set_error("NCCL transport requested but libnccl.so.2 cannot be loaded: %s", dlerror());
      delete t;
      return nullptr;
    }
    auto sym = [&](const char* name) { return dlsym(t->nccl_lib, name); };
    typedef int (*init_fn)(void**, int, NcclUniqueId, int);
    init_fn ncclCommInitRank = (init_fn)sym("ncclCommInitRank");
    t->ncclSend = (decltype(t->ncclSend))sym("ncclSend");
    t->ncclRecv = (decltype(t->ncclRecv))sym("ncclRecv");
    t->ncclGroupStart = (decltype(t->ncclGroupStart))sym("ncclGroupStart");
    t->ncclGroupEnd = (decltype(t->ncclGroupEnd))sym("ncclGroupEnd");
    t->ncclCommDestroy = (decltype(t->ncclCommDestroy))sym("ncclCommDestroy");
    t->ncclGetErrorString = (decltype(t->ncclGetErrorString))sym("ncclGetErrorString");
    if (!ncclCommInitRank || !t->ncclSend || !t->ncclRecv || !t->ncclGroupStart || !t->ncclGroupEnd) {
      set_error("libnccl is missing required symbols");
      delete t;
      return nullptr;
    }
    if (!c->cfg.nccl_unique_id) {
      set_error("NCCL transport needs cfg.nccl_unique_id (128 bytes, same on every rank)");
      delete t;
      return nullptr;
    }
    NcclUniqueId id;
    std::memcpy(&id, c->cfg.nccl_unique_id, sizeof(id));
    int r = ncclCommInitRank(&t->comm, c->cfg.nranks, id, c->cfg.rank);
    if (r != 0) {
      set_error("ncclCommInitRank failed: %s", t->ncclGetErrorString ? t->ncclGetErrorString(r) : "?");
      delete t;
      return nullptr;
    }
    // The particle exchange of a push runs on the side stream while the field phases exchange halos on the library
    // stream: two streams need two communicators (every rank splits: collective)
    typedef int (*split_fn)(void*, int, int, void**, void*);
    split_fn ncclCommSplit = (split_fn)sym("ncclCommSplit");
    if (ncclCommSplit && ncclCommSplit(t->comm, 0, c->cfg.rank, &t->comm2, nullptr) != 0) t->comm2 = nullptr;
  } else if (t->kind == CYLGPU_TRANSPORT_CALLBACK) {
    if (!c->cfg.sendrecv) {
      set_error("CALLBACK transport needs cfg.sendrecv");
      delete t;
      return nullptr;
    }
  } else if (t->kind == CYLGPU_TRANSPORT_FABRIC) {
    if (!c->cfg.fabric || ((Fabric*)c->cfg.fabric)->n != c->cfg.nranks) {
      set_error("FABRIC transport needs cfg.fabric created for nranks handles");
      delete t;
      return nullptr;
    }
  }
  return t;
}

void destroy_transport(Transport* t) {
  if (!t) return;
  p2p_release(t->p2p);
  if (t->comm2 && t->ncclCommDestroy) t->ncclCommDestroy(t->comm2);
  if (t->comm && t->ncclCommDestroy) t->ncclCommDestroy(t->comm);
  delete t;
}

// send `sl` to the left neighbour and `sr` to the right one, receive `rl` from the left and
// `rr` from the right.  Null / zero-byte legs are skipped (MPI_PROC_NULL).
int transport_sendrecv(cylgpu_ctx* c, const void* sl, size_t sl_b, void* rl, size_t rl_b, const void* sr,
                       size_t sr_b, void* rr, size_t rr_b) {
  Transport* t = c->tr;
  const int left = c->left, right = c->right;
  if (left < 0) { sl = nullptr; sl_b = 0; rl = nullptr; rl_b = 0; }
  if (right < 0) { sr = nullptr; sr_b = 0; rr = nullptr; rr_b = 0; }
  const bool self_l = (left == c->cfg.rank), self_r = (right == c->cfg.rank);
  if (self_l || self_r) {
    // periodic wrap onto myself: what I send left arrives as "from my right", and vice versa
    if (!(self_l && self_r)) { set_error("internal: half self-neighbour"); return 4; }
    if (sl_b != rr_b || sr_b != rl_b) { set_error("internal: self exchange size mismatch"); return 4; }
    if (sl_b) CUDA_TRY(cudaMemcpyAsync(rr, sl, sl_b, cudaMemcpyDeviceToDevice, c->stream));
    if (sr_b) CUDA_TRY(cudaMemcpyAsync(rl, sr, sr_b, cudaMemcpyDeviceToDevice, c->stream));
    return 0;
  }
  if (left < 0 && right < 0) return 0;
  switch (t->kind) {
    case CYLGPU_TRANSPORT_NCCL: {
      PhaseTimer timer(c, &c->stats.ms_exchange);   // device time of the exchanges (includes waiting for the peers)
      // Each link on its own: peer-memory mailboxes where both ends mapped and the messages fit a slot (both ends
      // see the same sizes, so they take the same branch), ncclSend / ncclRecv otherwise.
      P2P& P = t->p2p;
      const bool on_side = c->side && c->stream == c->side;   // (the mailboxes belong to the library stream)
      void* comm = (on_side && t->comm2) ? t->comm2 : t->comm;
      const bool l_on = left >= 0 && (sl_b || rl_b), r_on = right >= 0 && (sr_b || rr_b);
      const bool p2p_msg = c->p2p_policy == 1 || (c->p2p_policy == 2 && c->msg_counted);
      const bool l_p2p = p2p_msg && !on_side && l_on && P.ready && c->p2p_link_l && sl_b <= P.cap && rl_b <= P.cap && sl_b % 8 == 0 && rl_b % 8 == 0;
      const bool r_p2p = p2p_msg && !on_side && r_on && P.ready && c->p2p_link_r && sr_b <= P.cap && rr_b <= P.cap && sr_b % 8 == 0 && rr_b % 8 == 0;
      if (l_p2p || r_p2p) {
        const size_t box = P2P_FLAGS_BYTES + 2 * P.cap;
        unsigned char* my_from_l = P.mine;
        unsigned char* my_from_r = P.mine + box;
        const int counted = c->msg_counted ? 1 : 0;
        const size_t big = std::max(std::max(l_p2p ? sl_b : 0, r_p2p ? sr_b : 0), std::max(l_p2p ? rl_b : 0, r_p2p ? rr_b : 0));
        // (a counted message usually carries a fraction of its capacity)
        const unsigned nb = (unsigned)std::min<size_t>(std::max<size_t>(big / (counted ? 64 : 16) / 1024, 1), 64);
        // sends first (nobody waits for me before I have written), then the arrivals.  My left-going message lands
        // in the left neighbour's from_right box; arrivals are acknowledged in the header that holds my acks about
        // that neighbour's messages.
        P2PLeg s0 = {}, s1 = {}, r0 = {}, r1 = {};
        if (l_p2p && sl_b) s0 = P2PLeg{(const double*)sl, nullptr, sl_b / 8, P.peer_l + box, P.seq + 0, (P2PBoxHeader*)my_from_l, P.done + 0};
        if (r_p2p && sr_b) s1 = P2PLeg{(const double*)sr, nullptr, sr_b / 8, P.peer_r, P.seq + 1, (P2PBoxHeader*)my_from_r, P.done + 1};
        if (r_p2p && rr_b) r0 = P2PLeg{nullptr, (double*)rr, rr_b / 8, my_from_r, P.seq + 3, (P2PBoxHeader*)P.peer_r, P.done + 2};
        if (l_p2p && rl_b) r1 = P2PLeg{nullptr, (double*)rl, rl_b / 8, my_from_l, P.seq + 2, (P2PBoxHeader*)(P.peer_l + box), P.done + 3};
        if (s0.n8 || s1.n8) k_p2p_send<<<dim3(nb, 2), 256, 0, c->stream>>>(s0, s1, P.cap, counted, P.status);
        if (r0.n8 || r1.n8) k_p2p_recv<<<dim3(nb, 2), 256, 0, c->stream>>>(r0, r1, P.cap, counted, P.status);
        c->stats.kernel_launches += ((s0.n8 || s1.n8) ? 1 : 0) + ((r0.n8 || r1.n8) ? 1 : 0);
        CUDA_TRY(cudaGetLastError());
      }
      const bool l_nccl = l_on && !l_p2p, r_nccl = r_on && !r_p2p;
      if (!l_nccl && !r_nccl) return 0;
      int r = t->ncclGroupStart();
      // ncclUint8 = 1
      // order matters when left == right (2 ranks, periodic): per peer NCCL matches sends
      // and receives in issue order, and my left-going message must land in the peer's
      // "from the right" buffer
      if (r == 0 && l_nccl && sl_b) r = t->ncclSend(sl, sl_b, 1, left, comm, c->stream);
      if (r == 0 && r_nccl && sr_b) r = t->ncclSend(sr, sr_b, 1, right, comm, c->stream);
      if (r == 0 && r_nccl && rr_b) r = t->ncclRecv(rr, rr_b, 1, right, comm, c->stream);
      if (r == 0 && l_nccl && rl_b) r = t->ncclRecv(rl, rl_b, 1, left, comm, c->stream);
      int r2 = t->ncclGroupEnd();
      if (r == 0) r = r2;
      if (r != 0) {
        set_error("NCCL exchange failed: %s", t->ncclGetErrorString ? t->ncclGetErrorString(r) : "?");
        return 5;
      }
      return 0;
    }
    case CYLGPU_TRANSPORT_CALLBACK: {
      int r = c->cfg.sendrecv(c->cfg.sendrecv_user, left, right, sl, sl_b, rl, rl_b, sr, sr_b, rr, rr_b,
                              (void*)c->stream);
      if (r != 0) { set_error("sendrecv callback returned %d", r); return 5; }
      return 0;
    }
    case CYLGPU_TRANSPORT_FABRIC: {
      Fabric* F = (Fabric*)c->cfg.fabric;
      CUDA_TRY(cudaStreamSynchronize(c->stream));   // my send buffers are complete
      F->slots[c->cfg.rank] = Fabric::Slot{sl, sl_b, sr, sr_b};
      F->barrier();
      // pull: my left neighbour's right-going message, my right neighbour's left-going one
      if (left >= 0 && rl_b) {
        const Fabric::Slot& s = F->slots[left];
        if (s.send_r_b != rl_b) { set_error("fabric: size mismatch from left"); return 5; }
        CUDA_TRY(cudaMemcpyAsync(rl, s.send_r, rl_b, cudaMemcpyDeviceToDevice, c->stream));
      }
      if (right >= 0 && rr_b) {
        const Fabric::Slot& s = F->slots[right];
        if (s.send_l_b != rr_b) { set_error("fabric: size mismatch from right"); return 5; }
        CUDA_TRY(cudaMemcpyAsync(rr, s.send_l, rr_b, cudaMemcpyDeviceToDevice, c->stream));
      }
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      F->barrier();   // nobody reuses a send buffer before every pull has finished
      return 0;
    }
    default:
      set_error("no transport for a multi-rank exchange");
      return 5;
  }
}

// may an exchange be enqueued on the side stream while others run on the library stream?
bool transport_two_streams(const cylgpu_ctx* c) {
  const Transport* t = c->tr;
  if (!t) return false;
  if (c->left < 0 && c->right < 0) return true;
  if (c->left == c->cfg.rank) return true;                       // periodic wrap onto myself: device copies
  if (t->kind == CYLGPU_TRANSPORT_NCCL) return t->comm2 != nullptr;
  return false;   // callback / fabric transports block the host: no point
}

// a receive that timed out left a mark instead of hanging the device
int p2p_check(cylgpu_ctx* c) {
  Transport* t = c->tr;
  if (!t || !t->p2p.status) return 0;
  int st = 0;
  CUDA_TRY(cudaMemcpyAsync(&st, t->p2p.status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (st != 0) {
    set_error("neighbour exchange timed out (%s): a neighbour rank did not take part in the same sequence of calls",
              st == 1 ? "mailbox slot never acknowledged" : "message never arrived");
    return 5;
  }
  return 0;
}

}  // namespace cylgpu

extern "C" {

void* cylgpu_fabric_create(int nranks) { return new cylgpu::Fabric(nranks); }
void cylgpu_fabric_destroy(void* f) { delete (cylgpu::Fabric*)f; }

int cylgpu_nccl_unique_id(void* out128) {
  void* h = cylgpu::open_nccl();
  if (!h) { cylgpu::set_error("libnccl.so.2 cannot be loaded: %s", dlerror()); return 1; }
  typedef int (*fn)(cylgpu::NcclUniqueId*);
  fn f = (fn)dlsym(h, "ncclGetUniqueId");
  if (!f) { cylgpu::set_error("ncclGetUniqueId not found"); return 1; }
  cylgpu::NcclUniqueId id;
  int r = f(&id);
  if (r != 0) { cylgpu::set_error("ncclGetUniqueId failed (%d)", r); return 1; }
  std::memcpy(out128, &id, 128);
  return 0;
}

}  // extern "C"
