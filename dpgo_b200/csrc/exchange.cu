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

int dpgo_use_neighbor_poses(dpgo_handle h, int aux) {
  if (!h) { set_error("dpgo_use_neighbor_poses: null handle"); return DPGO_EINVAL; }
  double *buf = nullptr;
  const int rc = dpgo_neighbor_buffer(h, aux, &buf);
  if (rc != DPGO_OK) return rc;
  return dpgo_set_neighbor_poses_dev(h, buf);
}

}  // extern "C"
