// Public-pose exchange between agents inside the C-ABI: what PGOAgent::getSharedPoseDict /
// getAuxSharedPoseDict (ref: src/PGOAgent.cpp:97-146) hand out and updateNeighborPoses /
// updateAuxNeighborPoses (:650-702) take in -- driven by examples/MultiRobotExample.cpp:183-204 -- as
// packed device tiles moved by NCCL send/recv (NVLink) between ranks, or gathered straight into the
// receiver's neighbour buffer when both agents live on the same device.  One call = all messages of a
// round, one NCCL group, everything queued on the rank's stream: no host code between the solve that
// produces the poses and the send that publishes them.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: in a torchrun process that is the library torch has
// already loaded), so the product library keeps no link-time dependency on it and single-GPU users never
// page it in.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every function is resolved with dlsym
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "device_state.h"

namespace dpgo {

namespace {

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::once_flag g_nccl_once;
bool g_nccl_ok = false;

template <typename F>
bool bind(F &fn, const char *name) {
  fn = reinterpret_cast<F>(dlsym(g_nccl.lib, name));
  return fn != nullptr;
}

bool load_nccl() {
  std::call_once(g_nccl_once, []() {
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
      g_nccl.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return;
    g_nccl_ok = bind(g_nccl.GetUniqueId, "ncclGetUniqueId") && bind(g_nccl.CommInitRank, "ncclCommInitRank") &&
                bind(g_nccl.CommDestroy, "ncclCommDestroy") && bind(g_nccl.Send, "ncclSend") &&
                bind(g_nccl.Recv, "ncclRecv") && bind(g_nccl.GroupStart, "ncclGroupStart") &&
                bind(g_nccl.GroupEnd, "ncclGroupEnd") && bind(g_nccl.GetErrorString, "ncclGetErrorString");
  });
  if (!g_nccl_ok) set_error("NCCL is not available (dlopen libnccl.so.2: %s)", g_nccl.lib ? "missing symbols" : dlerror());
  return g_nccl_ok;
}

__global__ void k_gather_tiles_x(const double *slot, const int *idx, int num, int tile, double *out) {
  const size_t total = (size_t)num * tile;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / tile), q = (int)(t % tile);
    out[t] = slot[(size_t)idx[k] * tile + q];
  }
}

}  // namespace

#define NCCL_TRY(expr)                                                                             \
  do {                                                                                             \
    ncclResult_t _r = (expr);                                                                      \
    if (_r != ncclSuccess) {                                                                       \
      dpgo::set_error("%s:%d NCCL error in %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return DPGO_ECUDA;                                                                           \
    }                                                                                              \
  } while (0)
#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      dpgo::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return DPGO_ECUDA;                                                                           \
    }                                                                                              \
  } while (0)

}  // namespace dpgo

struct dpgo_comm_s {
  int device = 0, rank = 0, world = 1;
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;      // null when world == 1 (only same-device messages are possible)
  double *staging = nullptr;      // packed tiles of the remote sends of one exchange
  size_t staging_doubles = 0;
  int64_t exchanges = 0, launches = 0;
};

using namespace dpgo;

extern "C" {

int dpgo_comm_unique_id(unsigned char *id) {
  if (!id) { set_error("dpgo_comm_unique_id: null id"); return DPGO_EINVAL; }
  if (!load_nccl()) return DPGO_ECUDA;
  static_assert(sizeof(ncclUniqueId) == DPGO_COMM_ID_BYTES, "DPGO_COMM_ID_BYTES is NCCL's unique id size");
  ncclUniqueId u;
  NCCL_TRY(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return DPGO_OK;
}

int dpgo_comm_create(int device, int rank, int world, const unsigned char *id, void *stream, dpgo_comm *out) {
  if (!out || world < 1 || rank < 0 || rank >= world) { set_error("dpgo_comm_create: bad arguments"); return DPGO_EINVAL; }
  CUDA_TRY(cudaSetDevice(device));
  dpgo_comm_s *c = new dpgo_comm_s();
  c->device = device; c->rank = rank; c->world = world;
  c->stream = (cudaStream_t)stream;
  if (world > 1) {
    if (!id) { delete c; set_error("dpgo_comm_create: a unique id is required when world > 1"); return DPGO_EINVAL; }
    if (!load_nccl()) { delete c; return DPGO_ECUDA; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
      set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      delete c;
      return DPGO_ECUDA;
    }
  }
  *out = c;
  return DPGO_OK;
}

int dpgo_comm_destroy(dpgo_comm c) {
  if (!c) return DPGO_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->staging) cudaFree(c->staging);
  delete c;
  return DPGO_OK;
}

int dpgo_comm_launch_count(dpgo_comm c, int64_t *n) {
  if (!c || !n) { set_error("dpgo_comm_launch_count: null argument"); return DPGO_EINVAL; }
  *n = c->launches;
  return DPGO_OK;
}

int dpgo_neighbor_buffer(dpgo_handle h, int aux, double **dev_ptr) {
  if (!h || !dev_ptr || aux < 0 || aux > 1) { set_error("dpgo_neighbor_buffer: bad arguments"); return DPGO_EINVAL; }
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->d_nbr_xy[aux]) {
    const size_t bytes = (size_t)std::max(h->num_nbr_slots, 1) * h->r * (h->d + 1) * sizeof(double);
    CUDA_TRY(cudaMalloc((void **)&h->d_nbr_xy[aux], bytes));
    CUDA_TRY(cudaMemsetAsync(h->d_nbr_xy[aux], 0, bytes, h->stream));
  }
  *dev_ptr = h->d_nbr_xy[aux];
  return DPGO_OK;
}

int dpgo_exchange(dpgo_comm c, const dpgo_message *msgs, int n) {
  if (!c || (n > 0 && !msgs) || n < 0) { set_error("dpgo_exchange: bad arguments"); return DPGO_EINVAL; }
  CUDA_TRY(cudaSetDevice(c->device));
  // ---- validate, size the staging area of the remote sends
  size_t need = 0;
  for (int i = 0; i < n; ++i) {
    const dpgo_message &m = msgs[i];
    if (!m.src && !m.dst) { set_error("dpgo_exchange: message %d has neither a local sender nor a local receiver", i); return DPGO_EINVAL; }
    if (m.count < 0 || m.aux < 0 || m.aux > 1) { set_error("dpgo_exchange: message %d is malformed", i); return DPGO_EINVAL; }
    for (dpgo_handle h : {m.src, m.dst})
      if (h && (h->device != c->device || h->stream != c->stream)) {
        set_error("dpgo_exchange: message %d uses a handle of another device or stream than the communicator", i);
        return DPGO_EINVAL;
      }
    if (m.src && (m.slot < 0 || m.slot >= 4 || (m.count > 0 && !m.d_frames))) { set_error("dpgo_exchange: message %d has no frame list", i); return DPGO_EINVAL; }
    if (m.dst && (m.dst_offset < 0 || m.dst_offset + m.count > m.dst->num_nbr_slots)) {
      set_error("dpgo_exchange: message %d does not fit the receiver's neighbour slots", i);
      return DPGO_EINVAL;
    }
    if ((!m.src || !m.dst) && (c->world < 2 || m.peer < 0 || m.peer >= c->world || m.peer == c->rank)) {
      set_error("dpgo_exchange: message %d names an invalid peer rank %d", i, m.peer);
      return DPGO_EINVAL;
    }
    if (m.src && !m.dst) need += (size_t)m.count * m.src->r * (m.src->d + 1);
  }
  if (need > c->staging_doubles) {
    // the previous area may still be read by queued sends: release it in stream order
    if (c->staging) CUDA_TRY(cudaFreeAsync(c->staging, c->stream));
    c->staging = nullptr;
    CUDA_TRY(cudaMallocAsync((void **)&c->staging, need * sizeof(double), c->stream));
    c->staging_doubles = need;
  }
  // ---- pack: same-device messages land in the receiver's buffer, remote ones in the staging area
  size_t off = 0;
  std::vector<size_t> send_off(n, 0);
  for (int i = 0; i < n; ++i) {
    const dpgo_message &m = msgs[i];
    if (!m.src || m.count == 0) continue;
    const int tile = m.src->r * (m.src->d + 1);
    double *out;
    if (m.dst) {
      double *buf = nullptr;
      const int rc = dpgo_neighbor_buffer(m.dst, m.aux, &buf);
      if (rc != DPGO_OK) return rc;
      out = buf + (size_t)m.dst_offset * tile;
    } else {
      out = c->staging + off;
      send_off[i] = off;
      off += (size_t)m.count * tile;
    }
    const size_t total = (size_t)m.count * tile;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)m.src->num_sms * 4));
    k_gather_tiles_x<<<grid, 256, 0, c->stream>>>(m.src->d_slot[m.slot], m.d_frames, m.count, tile, out);
    c->launches++;
  }
  CUDA_TRY(cudaPeekAtLastError());
  // ---- one NCCL group: sends from the staging area, receives straight into the neighbour buffers.
  // Messages between one pair of ranks are matched in the order both sides list them.
  bool any = false;
  for (int i = 0; i < n; ++i) any = any || ((!msgs[i].src || !msgs[i].dst) && msgs[i].count > 0);
  if (any) {
    NCCL_TRY(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
      const dpgo_message &m = msgs[i];
      if (m.count == 0 || (m.src && m.dst)) continue;
      if (m.src) {
        const int tile = m.src->r * (m.src->d + 1);
        NCCL_TRY(g_nccl.Send(c->staging + send_off[i], (size_t)m.count * tile, ncclDouble, m.peer, c->comm, c->stream));
      } else {
        const int tile = m.dst->r * (m.dst->d + 1);
        double *buf = nullptr;
        const int rc = dpgo_neighbor_buffer(m.dst, m.aux, &buf);
        if (rc != DPGO_OK) { g_nccl.GroupEnd(); return rc; }
        NCCL_TRY(g_nccl.Recv(buf + (size_t)m.dst_offset * tile, (size_t)m.count * tile, ncclDouble, m.peer, c->comm, c->stream));
      }
    }
    NCCL_TRY(g_nccl.GroupEnd());
    c->launches++;
  }
  c->exchanges++;
  return DPGO_OK;
}

// ---- asynchronous publication through peer memory ----------------------------------------------------------------
// The reference's asynchronous mode (src/PGOAgent.cpp:475-499) lets every agent iterate at its own rate with whatever
// neighbour poses have arrived.  Here an agent PUBLISHES its public poses by storing them straight into a mailbox that
// lives in the memory of the neighbour's GPU (CUDA IPC mapping, NVLink peer stores), and COLLECTS consistent snapshots
// of its own mailboxes before a solve -- no rendezvous, no collective, no host in between.
//   mailbox = sequence word + 3 payload slots; message k goes to slot k % 3; the writer stores the payload, fences
//   (system scope) and then stores k; the reader loads k, copies slot k % 3 and loads the word again: the snapshot is
//   torn only if the writer has meanwhile started message k + 3, which it can do only after publishing k + 2 -- so a
//   second reading below k + 2 proves the copy consistent, anything else retries.
struct dpgo_mailbox_s {
  int device = 0;
  int count = 0, tile = 0;          // tiles per message, doubles per tile
  int dst_offset = 0;               // first neighbour slot of the owner the payload goes to
  dpgo_handle owner = nullptr;      // receiving agent (null on the sender's mapping)
  unsigned long long *base = nullptr;   // [16 words header | 3 payload slots]
  bool mapped = false;              // base comes from cudaIpcOpenMemHandle
  unsigned long long next_seq = 1;  // sender side: sequence number of the next message
};

namespace dpgo_mailbox_detail {
constexpr int kMailboxHeaderWords = 16;   // 128 bytes: the payload stays 128-byte aligned

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// payload of message `seq`: tiles frames[0..count) of `slot` -> slot seq % 3 of the (remote) mailbox
__global__ void k_mailbox_store(const double *slot, const int *frames, int count, int tile, unsigned long long *mb,
                                unsigned long long seq) {
  double *dst = reinterpret_cast<double *>(mb + kMailboxHeaderWords) + (size_t)(seq % 3) * count * tile;
  const size_t total = (size_t)count * tile;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(t / tile), q = (int)(t % tile);
    dst[t] = slot[(size_t)frames[k] * tile + q];
  }
}
// publication: runs after k_mailbox_store on the same stream (all payload stores are complete), makes them visible
// system-wide, then releases the sequence number
__global__ void k_mailbox_release(unsigned long long *mb, unsigned long long seq) {
  __threadfence_system();
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(mb), "l"(seq) : "memory");
}

struct MailboxView {
  const unsigned long long *mb;
  double *dst;          // where the snapshot goes (receiver's neighbour buffer)
  int doubles;          // count * tile
};
// one CTA per mailbox: consistent snapshot of the newest message (nothing is copied before the first message)
__global__ void k_mailbox_collect(const MailboxView *views) {
  const MailboxView v = views[blockIdx.x];
  __shared__ unsigned long long s_seq;
  for (int attempt = 0; attempt < 64; ++attempt) {
    if (threadIdx.x == 0) s_seq = ld_acquire_sys(v.mb);
    __syncthreads();
    const unsigned long long seq = s_seq;
    if (seq == 0) return;                                   // nothing published yet: keep what is there
    const double *src = reinterpret_cast<const double *>(v.mb + kMailboxHeaderWords) + (size_t)(seq % 3) * v.doubles;
    for (int t = threadIdx.x; t < v.doubles; t += blockDim.x) {
      double x;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(x) : "l"(src + t) : "memory");
      v.dst[t] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_seq = ld_acquire_sys(v.mb);
    __syncthreads();
    if (s_seq < seq + 2) return;                            // the writer cannot have touched slot seq % 3
    __syncthreads();
  }
}
}  // namespace dpgo_mailbox_detail
using namespace dpgo_mailbox_detail;

int dpgo_mailbox_create(dpgo_handle h, int dst_offset, int count, dpgo_mailbox *out, unsigned char *ipc_handle) {
  if (!h || !out || count < 0 || dst_offset < 0 || dst_offset + count > h->num_nbr_slots) {
    set_error("dpgo_mailbox_create: bad arguments");
    return DPGO_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  dpgo_mailbox_s *m = new dpgo_mailbox_s();
  m->device = h->device; m->count = count; m->tile = h->r * (h->d + 1); m->dst_offset = dst_offset; m->owner = h;
  const size_t bytes = kMailboxHeaderWords * 8 + (size_t)3 * std::max(count, 1) * m->tile * sizeof(double);
  if (cudaMalloc((void **)&m->base, bytes) != cudaSuccess || cudaMemset(m->base, 0, bytes) != cudaSuccess) {
    set_error("dpgo_mailbox_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete m;
    return DPGO_ECUDA;
  }
  if (ipc_handle) {
    static_assert(sizeof(cudaIpcMemHandle_t) == DPGO_IPC_HANDLE_BYTES, "DPGO_IPC_HANDLE_BYTES is CUDA's IPC handle size");
    cudaIpcMemHandle_t ih;
    if (cudaIpcGetMemHandle(&ih, m->base) != cudaSuccess) {
      set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
      cudaFree(m->base);
      delete m;
      return DPGO_ECUDA;
    }
    memcpy(ipc_handle, &ih, sizeof(ih));
  }
  h->mailboxes.push_back(m);
  h->mailbox_views_dirty = true;
  *out = m;
  return DPGO_OK;
}

int dpgo_mailbox_open(int device, const unsigned char *ipc_handle, dpgo_mailbox local, int count, int tile,
                      dpgo_mailbox *out) {
  if (!out || (!ipc_handle && !local) || count < 0 || tile <= 0) { set_error("dpgo_mailbox_open: bad arguments"); return DPGO_EINVAL; }
  CUDA_TRY(cudaSetDevice(device));
  dpgo_mailbox_s *m = new dpgo_mailbox_s();
  m->device = device; m->count = count; m->tile = tile;
  if (local) {                       // the receiver lives in this process: its mailbox is used directly
    if (local->count != count || local->tile != tile) { delete m; set_error("dpgo_mailbox_open: shape mismatch"); return DPGO_EINVAL; }
    m->base = local->base;
  } else {
    cudaIpcMemHandle_t ih;
    memcpy(&ih, ipc_handle, sizeof(ih));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, ih, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
      delete m;
      return DPGO_ECUDA;
    }
    m->base = (unsigned long long *)p;
    m->mapped = true;
  }
  *out = m;
  return DPGO_OK;
}

int dpgo_mailbox_close(dpgo_mailbox m) {
  if (!m) return DPGO_OK;
  cudaSetDevice(m->device);
  if (m->owner) {
    auto &v = m->owner->mailboxes;
    v.erase(std::remove(v.begin(), v.end(), m), v.end());
    m->owner->mailbox_views_dirty = true;
    cudaFree(m->base);
  } else if (m->mapped) {
    cudaIpcCloseMemHandle(m->base);
  }
  delete m;
  return DPGO_OK;
}

int dpgo_publish(dpgo_handle h, int slot, dpgo_mailbox to, const int32_t *d_frames) {
  if (!h || !to || to->owner || slot < 0 || slot >= 4 || (to->count > 0 && !d_frames) || to->tile != h->r * (h->d + 1)) {
    set_error("dpgo_publish: bad arguments (the mailbox must be a sender-side mapping of matching tile size)");
    return DPGO_EINVAL;
  }
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned long long seq = to->next_seq++;
  if (to->count > 0) {
    const size_t total = (size_t)to->count * to->tile;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 2));
    k_mailbox_store<<<grid, 256, 0, h->stream>>>(h->d_slot[slot], d_frames, to->count, to->tile, to->base, seq);
  }
  k_mailbox_release<<<1, 1, 0, h->stream>>>(to->base, seq);
  h->launches += 2;
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

int dpgo_collect(dpgo_handle h) {
  if (!h) { set_error("dpgo_collect: null handle"); return DPGO_EINVAL; }
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->mailboxes.empty()) return DPGO_OK;
  double *buf = nullptr;
  const int rc = dpgo_neighbor_buffer(h, 0, &buf);
  if (rc != DPGO_OK) return rc;
  if (h->mailbox_views_dirty) {
    std::vector<MailboxView> v;
    for (dpgo_mailbox m : h->mailboxes)
      v.push_back(MailboxView{m->base, buf + (size_t)m->dst_offset * m->tile, m->count * m->tile});
    if (h->d_mailbox_views) CUDA_TRY(cudaFree(h->d_mailbox_views));
    CUDA_TRY(cudaMalloc(&h->d_mailbox_views, v.size() * sizeof(MailboxView)));
    CUDA_TRY(cudaMemcpyAsync(h->d_mailbox_views, v.data(), v.size() * sizeof(MailboxView), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));   // v goes out of scope
    h->mailbox_views_dirty = false;
  }
  k_mailbox_collect<<<(int)h->mailboxes.size(), 256, 0, h->stream>>>((const MailboxView *)h->d_mailbox_views);
  h->launches++;
  CUDA_TRY(cudaPeekAtLastError());
  return DPGO_OK;
}

int dpgo_use_neighbor_poses(dpgo_handle h, int aux) {
  if (!h) { set_error("dpgo_use_neighbor_poses: null handle"); return DPGO_EINVAL; }
  double *buf = nullptr;
  const int rc = dpgo_neighbor_buffer(h, aux, &buf);
  if (rc != DPGO_OK) return rc;
  return dpgo_set_neighbor_poses_dev(h, buf);
}

}  // extern "C"
