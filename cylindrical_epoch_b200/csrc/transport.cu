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

#include <condition_variable>
#include <cstring>
#include <mutex>

#include "ctx.cuh"

namespace cylgpu {

struct Transport {
  int kind = CYLGPU_TRANSPORT_NONE;
  // NCCL
  void* nccl_lib = nullptr;
  void* comm = nullptr;
  int (*ncclSend)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ncclRecv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ncclGroupStart)() = nullptr;
  int (*ncclGroupEnd)() = nullptr;
  int (*ncclCommDestroy)(void*) = nullptr;
  const char* (*ncclGetErrorString)(int) = nullptr;
};

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
      int r = t->ncclGroupStart();
      // ncclUint8 = 1
      // order matters when left == right (2 ranks, periodic): per peer NCCL matches sends
      // and receives in issue order, and my left-going message must land in the peer's
      // "from the right" buffer
      if (r == 0 && sl_b) r = t->ncclSend(sl, sl_b, 1, left, t->comm, c->stream);
      if (r == 0 && sr_b) r = t->ncclSend(sr, sr_b, 1, right, t->comm, c->stream);
      if (r == 0 && rr_b) r = t->ncclRecv(rr, rr_b, 1, right, t->comm, c->stream);
      if (r == 0 && rl_b) r = t->ncclRecv(rl, rl_b, 1, left, t->comm, c->stream);
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
