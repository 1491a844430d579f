// Block-CSR Q*X for problems that stream from HBM: shared-memory staging of the Q block runs and of the
// gathered r x (d+1) pose tiles with asynchronous copies (cp.async, SASS LDGSTS), one block ROW per warp.
//   out = X * Q (+ G)     ref: QuadraticProblem::f / EucGrad / EucHessianEta, src/QuadraticProblem.cpp:29-54
//
// Why not the lane-group kernel (phase_qx) at scale: there a group of d+1 lanes walks its block row
// sequentially, every step a dependent chain colidx -> X tile -> FMA with one block's loads in flight, and each
// lane loads the whole 160-byte tile (10 x LDG.128 whose 32 lanes touch 8 different lines).  Here
//   * lanes are (block slot b, column c): a warp takes GPW = 32/(d+1) blocks of one row per step, so all the
//     tiles of a row are gathered at once;
//   * the gathers are 16-byte cp.async pieces dealt over the lanes (3 tiles per warp instruction, whole lines),
//     the Q blocks of the step are one contiguous run copied the same way; nothing is held in registers while in
//     flight, and a warp keeps kQxDepth steps in flight ahead of the one it computes (column indices are fetched
//     one step further ahead, the row pointers 16 rows ahead), so the DRAM stream of Q never waits for a gather;
//   * the compute step reads tile and block row from shared memory (broadcast inside a lane group), and the
//     partial sums of the block slots are combined by a fixed shuffle tree (deterministic).
// Rows are dealt round-robin over all resident warps so that the warps sweep the matrix together and the X tiles
// they gather stay inside a sliding window of the L2.
#pragma once
#include "kernels.cuh"

namespace dpgo {

constexpr int kQxDepth = 2;              // steps in flight ahead of the computed one
constexpr int kQxBufs = kQxDepth + 1;
constexpr int kQxIdxAhead = 2;           // steps whose column indices are loaded ahead of their copies
constexpr int kQxWindow = kQxDepth + 1 + kQxIdxAhead;

template <int R, int D>
struct QxGeo {
  static constexpr int DH = D + 1;
  static constexpr int GPW = 32 / DH;                          // block slots per warp
  static constexpr int TILE = R * DH;                          // doubles per pose tile
  static constexpr int BLK = DH * DH;                          // doubles per Q block
  static constexpr int PB = (TILE % 2 == 0) ? 16 : 8;          // bytes per gather piece
  static constexpr int NP = TILE * 8 / PB;                     // pieces per tile
  static constexpr int PBQ = (BLK % 2 == 0) ? 16 : 8;
  static constexpr int NPQ = BLK * 8 / PBQ;                    // pieces per block
  // tile pitch in shared memory: + 16 bytes so that the GPW tiles a warp instruction reads start in different banks
  static constexpr int TP = TILE * 8 + 16;
  static constexpr int QBYTES = GPW * BLK * 8;
  static constexpr int STEP_BYTES = ((QBYTES + GPW * TP + 127) / 128) * 128;
  static constexpr int WARP_BYTES = kQxBufs * STEP_BYTES;
  static constexpr int CTA_BYTES = kWarpsPerBlock * WARP_BYTES;
};

#ifndef DPGO_CPU_EMU
// STREAM: data read once (the Q blocks) -- 16-byte copies bypass the L1 (.cg); 8-byte copies exist only as .ca
template <int BYTES, bool STREAM = false>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  if constexpr (STREAM && BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// what a warp knows about one step (warp-uniform except j)
struct QxStep {
  int row;     // block row; < 0: no more work
  int e;       // first block of the step
  int cnt;     // blocks in the step (<= GPW)
  int last;    // last step of its row
  int first;   // first step of its row
  int j;       // column index of this lane's block slot
};

template <int R, int D>
__device__ __forceinline__ void phase_qx_staged(const BsrView &Q, const double *X, const double *G, double *out,
                                                int n, unsigned char *smem, int warp_global, int nwarps_global) {
  using Gm = QxGeo<R, D>;
  constexpr int DH = Gm::DH, GPW = Gm::GPW;
  const int lane = threadIdx.x & 31;
  const int b = lane / DH, c = lane - b * DH;
  const bool ok = b < GPW;
  unsigned char *wbuf = smem + (size_t)(threadIdx.x >> 5) * Gm::WARP_BYTES;

  // ---- row pointers of the next 16 rows of this warp (lanes 0..15: begin, 16..31: end)
  int meta = 0, meta_k0 = 0, meta_n = 0;   // rows warp_global + (meta_k0 + i) * nwarps_global, i < meta_n
  int k_next = 0;                          // index of the next row of this warp to start
  int cur_e = 0, cur_e1 = 0, cur_row = -1;
  auto refill_meta = [&]() {
    meta_k0 = k_next;
    const long row = (long)warp_global + (long)(meta_k0 + (lane & 15)) * nwarps_global;
    meta = (row < n) ? __ldg(Q.rowptr + row + (lane >> 4)) : -1;
    meta_n = 16;
  };
  // next step of this warp's sequence, with its column index load issued
  auto generate = [&]() -> QxStep {
    QxStep s;
    s.first = 0;
    if (cur_e >= cur_e1) {                 // start the next row (empty rows still produce one step: out = G)
      if (k_next - meta_k0 >= meta_n) refill_meta();
      const int slot = k_next - meta_k0;
      const int ra = __shfl_sync(0xffffffffu, meta, slot), rb = __shfl_sync(0xffffffffu, meta, slot + 16);
      if (ra < 0) {
        s.row = -1; s.e = 0; s.cnt = 0; s.last = 0; s.j = 0;
        return s;
      }
      cur_row = warp_global + k_next * nwarps_global;
      cur_e = ra;
      cur_e1 = rb;
      k_next++;
      s.first = 1;
    }
    s.row = cur_row;
    s.e = cur_e;
    s.cnt = min(GPW, cur_e1 - cur_e);
    cur_e += s.cnt;
    s.last = cur_e >= cur_e1;
    s.j = (ok && b < s.cnt) ? __ldg(Q.colidx + s.e + b) : 0;
    return s;
  };
  // asynchronous copies of one step into buffer `buf`
  auto issue = [&](const QxStep &s, int buf) {
    if (s.row >= 0 && s.cnt > 0) {
      unsigned char *qb = wbuf + (size_t)buf * Gm::STEP_BYTES, *xb = qb + Gm::QBYTES;
      const unsigned char *qsrc = reinterpret_cast<const unsigned char *>(Q.blocks + (size_t)s.e * Gm::BLK);
      const int nq = s.cnt * Gm::NPQ;
      for (int p = lane; p < nq; p += 32) cp_async<Gm::PBQ, true>(qb + p * Gm::PBQ, qsrc + (size_t)p * Gm::PBQ);
      const int nx = s.cnt * Gm::NP;
      for (int p0 = 0; p0 < nx; p0 += 32) {          // uniform trip count: every lane takes part in the shuffle
        const int p = p0 + lane;
        const bool valid = p < nx;
        const int tile = valid ? p / Gm::NP : 0, piece = p - tile * Gm::NP;
        const int j = __shfl_sync(0xffffffffu, s.j, tile * DH);
        if (valid)
          cp_async<Gm::PB>(xb + tile * Gm::TP + piece * Gm::PB,
                           reinterpret_cast<const unsigned char *>(X + (size_t)j * Gm::TILE) + piece * Gm::PB);
      }
    }
    cp_async_commit();                                // one group per step, empty or not: the wait counts groups
  };

  // ---- pipeline: steps t+1 .. t+kQxDepth are in flight while step t is computed; the column indices of
  // steps up to t+kQxDepth+kQxIdxAhead are being loaded
  QxStep st[kQxWindow];
#pragma unroll
  for (int i = 0; i < kQxWindow; ++i) st[i] = generate();
#pragma unroll
  for (int i = 0; i <= kQxDepth; ++i) issue(st[i], i);    // steps 0 .. kQxDepth (uses their indices)
  double acc[R];
#pragma unroll
  for (int q = 0; q < R; ++q) acc[q] = 0.0;
  int buf = 0;
  while (st[0].row >= 0) {
    cp_async_wait<kQxDepth>();                            // all but the kQxDepth most recent groups have landed
    __syncwarp();
    const QxStep s = st[0];
    if (s.first) {
#pragma unroll
      for (int q = 0; q < R; ++q) acc[q] = 0.0;
    }
    if (ok && b < s.cnt) {
      const unsigned char *qb = wbuf + (size_t)buf * Gm::STEP_BYTES, *xb = qb + Gm::QBYTES;
      const double *m = reinterpret_cast<const double *>(qb) + b * Gm::BLK + c * DH;   // row c of the block
      const double *x = reinterpret_cast<const double *>(xb + b * Gm::TP);
      double mk[DH], xv[Gm::TILE];
      if constexpr (DH % 2 == 0) {
#pragma unroll
        for (int k = 0; k < DH / 2; ++k) {
          const double2 v = reinterpret_cast<const double2 *>(m)[k];
          mk[2 * k] = v.x; mk[2 * k + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < DH; ++k) mk[k] = m[k];
      }
      if constexpr (Gm::TILE % 2 == 0) {
#pragma unroll
        for (int k = 0; k < Gm::TILE / 2; ++k) {
          const double2 v = reinterpret_cast<const double2 *>(x)[k];
          xv[2 * k] = v.x; xv[2 * k + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < Gm::TILE; ++k) xv[k] = x[k];
      }
#pragma unroll
      for (int k = 0; k < DH; ++k)
#pragma unroll
        for (int q = 0; q < R; ++q) acc[q] = fma(xv[k * R + q], mk[k], acc[q]);
    }
    if (s.last) {
      // combine the block slots: fixed tree over b (slot b takes slot b + o), result in the lanes of slot 0
#pragma unroll
      for (int o = 8; o >= 1; o >>= 1) {
        if (o < GPW) {
#pragma unroll
          for (int q = 0; q < R; ++q) {
            const double v = __shfl_down_sync(0xffffffffu, acc[q], o * DH);
            if (b + o < GPW) acc[q] += v;
          }
        }
      }
      if (b == 0) {
        const size_t off = ((size_t)s.row * DH + c) * R;
        if (G) {
#pragma unroll
          for (int q = 0; q < R; ++q) acc[q] += G[off + q];
        }
        store_col<R>(out + off, acc);
      }
    }
    __syncwarp();                                         // every lane has read buffer `buf`: it may be refilled
    // shift the window, start the copies of the step that enters it
#pragma unroll
    for (int i = 0; i < kQxWindow - 1; ++i) st[i] = st[i + 1];
    st[kQxWindow - 1] = generate();
    issue(st[kQxDepth], buf);                             // step t + kQxDepth + 1 reuses the buffer just read
    buf = (buf + 1 == kQxBufs) ? 0 : buf + 1;
  }
  cp_async_wait<0>();
}

}  // namespace dpgo
