// Coarse cosine scoring of a query tile against a descriptor pool on the
// 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by
// TMA into 128B-swizzled shared memory), with the candidate filter fused into
// the epilogue.
//
// Replaces the per-row Python loop of NearestNeighborsMatching.search
// (reference cslam/nns_matching.py:55-58) for the *candidate generation*
// stage; the exact float64 scores the reference returns are recomputed for the
// survivors by k_nns_select_rerank (nns.cu).
//
//   D[128 queries, 256 pool rows] += Qh[128, 64] * Ph[256, 64]^T      (fp16 in, fp32 acc)
//
// Layout
//   * Qh  [128*T, dim_pad] fp16, rows L2-normalised, zero padded   (A operand, K-major)
//   * Ph  [cap,   dim_pad] fp16, rows L2-normalised, zero padded   (B operand, K-major)
//   * TMEM: 2 accumulator buffers x 256 fp32 columns = all 512 columns; TMEM lane m
//     holds query m, column j holds pool row (tile_row0 + j).
// Query batches wider than 128 run as thread-block CLUSTERS of C = 2 or 4 CTAs: CTA r of a
// cluster owns query tile r of the group, all CTAs of a cluster walk the same pool tiles, and
// every pool tile is fetched from HBM ONCE per cluster: each CTA loads 256/C of its rows and
// TMA-multicasts them into the shared memory of all C CTAs.  A 512-query batch then costs one
// pool sweep instead of four and the kernel moves from the HBM roofline to the tensor pipe.
// Warp roles (384 threads, 1 CTA / SM, persistent over pool tiles):
//   warp 0    TMA producer (one lane)       warp 1  MMA issuer (one lane)
//   warp 2    TMEM allocator                warp 3  idle
//   warps 4-11 epilogue: thread = (query, column half); tcgen05.ld 32 columns at a
//             time, compare with the query's threshold tau, append survivors
//             (score, row) to the thread's private sub-segment of the query's candidate
//             list in global memory (no atomics).  In sample mode only the maximum of
//             each 32-row chunk is stored (input of the threshold selection).
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "nns_internal.cuh"
#include "tc_common.cuh"

namespace cslam {
using namespace tc;
namespace {

constexpr int BM = kCoarseBM;  // queries per tile (UMMA M)
constexpr int BN = kCoarseBN;  // pool rows per tile (UMMA N)
constexpr int BK = 64;         // fp16 elements per k-block = one 128B swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
// "A-resident" mode (clusters, dim_pad <= 512): the query tile stays in shared memory for the
// whole kernel (num_kb x 16 KiB) and only pool tiles stream through a 3-stage ring, which cuts
// the bytes every SM has to ingest per MMA by a third (the chip-wide TMA/L2 delivery rate, not
// HBM, is what bounds the cluster kernel).
constexpr int RES_MAX_KB = 8;
constexpr int RES_STAGES = 3;
constexpr int RES_DATA_BYTES = RES_MAX_KB * A_BYTES + RES_STAGES * B_BYTES;   // 224 KiB
constexpr int SMEM_BYTES_RES = RES_DATA_BYTES + 1024 + 256;
static_assert(SMEM_BYTES_RES <= 232448, "A-resident layout exceeds the 227 KiB CTA limit");

// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=f16,
// both K-major, N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
                            (static_cast<uint32_t>(BN >> 3) << 17) |
                            (static_cast<uint32_t>(BM >> 4) << 24);

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int nst) {
    if (++stage == nst) { stage = 0; phase ^= 1; }
  }
};

// Epilogue work of one thread on one accumulator tile: its query's scores for one half of the
// tile's 256 pool rows, 32 columns per tcgen05.ld.  Sample mode stores chunk maxima, filter mode
// appends (score, row) pairs clearing the query's threshold to the thread's private segment.
__device__ __forceinline__ void scan_accumulator_half(
    const CoarseParams& prm, uint32_t taddr_buf, int half, int row0, int tile, float tau,
    bool q_valid, uint2* my_seg, uint2* my_ovf, unsigned int* my_cnt, uint32_t* my_smax,
    unsigned int& my_count) {
  constexpr int HALF_N = BN / 2;
#pragma unroll 1
      for (int c = half * HALF_N; c < (half + 1) * HALF_N; c += kChunk) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr_buf + static_cast<uint32_t>(c), r);
        tmem_ld_wait();
        float m[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          m[j] = fmaxf(fmaxf(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])),
                       fmaxf(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        const float mx = fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])),
                               fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
        if (prm.mode == 1) {
          if (q_valid) my_smax[tile * (BN / kChunk) + c / kChunk] = f32_to_key(mx);
        } else if (mx >= tau) {
          // Per-thread hit path (a thread owns its candidate segment, so no warp votes are
          // needed): only the 4-element groups whose maximum clears the threshold are
          // examined.  With the sampled threshold ~1000 rows per query clear tau, i.e. most
          // 32x32 chunks contain a hit for SOME lane - this path is hot and must stay short.
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (m[g] >= tau) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = 4 * g + e;
                const float s = __uint_as_float(r[j]);
                const int row = row0 + c + j;
                if (s >= tau && row < prm.n_rows) {
                  const uint2 ent = make_uint2(__float_as_uint(s), static_cast<uint32_t>(row));
                  if (my_count < static_cast<unsigned int>(kSegCap)) {
                    my_seg[my_count] = ent;
                  } else {
                    const unsigned int pos = atomicAdd(my_cnt + prm.nsub, 1u);
                    if (pos < static_cast<unsigned int>(kOvfCap)) my_ovf[pos] = ent;
                  }
                  ++my_count;
                }
              }
            }
          }
        }
      }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_nns_coarse_tc(const __grid_constant__ CUtensorMap tmap_q,
                const __grid_constant__ CUtensorMap tmap_p, CoarseParams prm) {
  // cluster geometry: C CTAs, rank r owns query tile r of the group; C == 1 is the plain kernel
  const int C = prm.cluster;
  const int crank = C > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int cid = C > 1 ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int ncl = C > 1 ? static_cast<int>(num_clusters_x()) : static_cast<int>(gridDim.x);
  const uint16_t cmask = static_cast<uint16_t>((1u << C) - 1u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const bool a_res = prm.a_resident != 0;
  const int nst = a_res ? RES_STAGES : STAGES;
  const uint32_t bar_base = smem_base + (a_res ? RES_DATA_BYTES : STAGES * STAGE_BYTES);
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  // streaming layout: stage = [A k-block | B k-block]; resident layout: [A k-blocks 0..7 | B ring]
  auto smem_a = [&](int s, int kb) {
    return a_res ? smem_base + kb * A_BYTES : smem_base + s * STAGE_BYTES;
  };
  auto smem_b = [&](int s) {
    return a_res ? smem_base + RES_MAX_KB * A_BYTES + s * B_BYTES
                 : smem_base + s * STAGE_BYTES + A_BYTES;
  };
  // barrier slots: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem slot, a_full
  const uint32_t afull_bar = bar_base + 8u * (2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_p) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), C);   // one MMA commit per CTA of the cluster frees a slot
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 256);
    }
    mbar_init(afull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int num_kb = prm.num_kb;
  long long* dbg = (prm.dbg != nullptr && blockIdx.x == 0) ? prm.dbg : nullptr;
  long long w_acc0 = 0, w_acc1 = 0;
  auto timed_wait = [&](uint32_t bar, uint32_t parity, long long& acc) {
    if (dbg) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const int first_tile = cid;
  const int tile_step = ncl;
  const int q_tile_row0 = prm.q_row0 + crank * BM;

  if (warp == 0) {
    if (lane == 0) {
      PipeState ps;
      if (a_res) {   // the whole query tile, once
        mbar_expect_tx(afull_bar, static_cast<uint32_t>(num_kb) * A_BYTES);
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_2d(smem_a(0, kb), &tmap_q, afull_bar, kb * BK, q_tile_row0, kEvictLast);
      }
      // The ring in shared memory holds well under one bandwidth-delay product of an HBM
      // stream, so in the cluster kernel (where every SM must ingest a whole pool tile per
      // 2.2 us of MMA) the tiles are pulled into L2 `pf` tiles ahead and the ring is refilled
      // at L2 latency.
      const int pf = prm.l2_prefetch;
      const int slice_rows = BN / C;
      for (int t = 0; t < pf; ++t) {
        const int tl = first_tile + t * tile_step;
        if (tl < prm.num_tiles) {
          const int r0 = static_cast<int>(static_cast<int64_t>(tl) * prm.tile_stride * BN);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_prefetch_l2_2d(&tmap_p, kb * BK, r0 + crank * slice_rows);
        }
      }
      for (int tile = first_tile; tile < prm.num_tiles; tile += tile_step) {
        const int row0 = static_cast<int>(static_cast<int64_t>(tile) * prm.tile_stride * BN);
        if (pf > 0) {
          const int tl = tile + pf * tile_step;
          if (tl < prm.num_tiles) {
            const int r0 = static_cast<int>(static_cast<int64_t>(tl) * prm.tile_stride * BN);
            for (int kb = 0; kb < num_kb; ++kb)
              tma_prefetch_l2_2d(&tmap_p, kb * BK, r0 + crank * slice_rows);
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          timed_wait(empty_bar(ps.stage), ps.phase ^ 1, w_acc0);
          mbar_expect_tx(full_bar(ps.stage), a_res ? B_BYTES : STAGE_BYTES);
          if (!a_res)
            tma_load_2d(smem_a(ps.stage, kb), &tmap_q, full_bar(ps.stage), kb * BK, q_tile_row0,
                        kEvictLast);
          if (C == 1) {
            tma_load_2d(smem_b(ps.stage), &tmap_p, full_bar(ps.stage), kb * BK, row0,
                        kEvictFirst);
          } else {
            // this CTA's slice of the pool tile, multicast into every CTA of the cluster
            const int rows = BN / C;
            tma_load_2d_mc(smem_b(ps.stage) + crank * rows * (BK * 2), &tmap_p, full_bar(ps.stage),
                           kb * BK, row0 + crank * rows, cmask, kEvictFirst);
          }
          ps.advance(nst);
        }
      }
      if (dbg) dbg[0] += w_acc0;   // producer: cycles waiting for a free smem slot
    }
  } else if (warp == 1) {
    if (lane == 0) {
      PipeState ps;
      int it = 0;
      if (a_res) {
        mbar_wait(afull_bar, 0);
        tc_fence_after();
      }
      for (int tile = first_tile; tile < prm.num_tiles; tile += tile_step, ++it) {
        const int buf = it & 1;
        timed_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, w_acc1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          timed_wait(full_bar(ps.stage), ps.phase, w_acc0);
          tc_fence_after();
          const uint32_t a0 = smem_a(ps.stage, kb);
          const uint32_t b0 = smem_b(ps.stage);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_sw128_desc(a0 + k * (UMMA_K * 2));
            const uint64_t bdesc = make_sw128_desc(b0 + k * (UMMA_K * 2));
            umma_f16(d_tmem, adesc, bdesc, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot once the MMAs retire (in every CTA that multicasts into it)
          if (C == 1) umma_commit(empty_bar(ps.stage));
          else umma_commit_mc(empty_bar(ps.stage), cmask);
          ps.advance(nst);
        }
        umma_commit(tfull_bar(buf));  // accumulator complete -> epilogue
      }
      if (dbg) {
        dbg[1] += w_acc0;   // MMA issuer: cycles waiting for operands
        dbg[2] += w_acc1;   // MMA issuer: cycles waiting for a free accumulator
      }
    }
  } else if (warp >= 4) {
    // 8 epilogue warps: warp % 4 selects the TMEM lane quarter (hardware rule), the
    // upper/lower group of four warps takes column half 0/1 of every tile.
    const int ew = warp - 4;
    const int quarter = ew & 3;
    const int half = ew >> 2;
    const int q = crank * BM + quarter * 32 + lane;   // query index within the group
    const bool q_valid = q < prm.nq;
    const int qs = q_valid ? q : 0;
    // invalid (padding) queries never hit: +inf threshold
    const float tau = !q_valid ? INFINITY : (prm.tau != nullptr ? prm.tau[q] : -INFINITY);
    const size_t slots = cand_slots(prm.nsub);
    uint2* my_seg = prm.cand + static_cast<size_t>(qs) * slots +
                    static_cast<size_t>(cid * 2 + half) * kSegCap;
    uint2* my_ovf = prm.cand + static_cast<size_t>(qs) * slots +
                    static_cast<size_t>(prm.nsub) * kSegCap;
    unsigned int* my_cnt = prm.cnt + static_cast<size_t>(qs) * (prm.nsub + 1);
    uint32_t* my_smax = prm.smax + static_cast<size_t>(qs) * prm.smax_stride;
    unsigned int my_count = 0;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int it = 0;
    for (int tile = first_tile; tile < prm.num_tiles; tile += tile_step, ++it) {
      const int buf = it & 1;
      const int row0 = static_cast<int>(static_cast<int64_t>(tile) * prm.tile_stride * BN);
      timed_wait(tfull_bar(buf), (it >> 1) & 1, w_acc0);
      tc_fence_after();
      const long long te0 = dbg ? clock64() : 0;
      scan_accumulator_half(prm, lane_taddr + static_cast<uint32_t>(buf * BN), half, row0, tile, tau,
                            q_valid, my_seg, my_ovf, my_cnt, my_smax, my_count);
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (dbg) w_acc1 += clock64() - te0;
    }
    if (dbg && threadIdx.x == 128) {
      dbg[3] += w_acc0;   // epilogue: cycles waiting for an accumulator
      dbg[4] += w_acc1;   // epilogue: cycles scanning accumulators
      dbg[5] += it;       // tiles
    }
    if (q_valid && prm.mode == 0)
      my_cnt[cid * 2 + half] = min(my_count, static_cast<unsigned int>(kSegCap));
  }

  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();   // nobody leaves while a peer may still write into its smem
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// SM-pair variant (tcgen05 cta_group::2) for query groups of 129..256: the two CTAs of a
// cluster form one MMA of M = 256 (each CTA's 128 queries) x N = 256 pool rows.  Each CTA stages
// only HALF of every pool tile (128 rows); the tensor cores of both SMs read both halves, so
// an SM ingests 128 KB instead of 256 KB per tile and the kernel is bound by HBM (one sweep per
// 256 queries) instead of by L2->SM delivery.  Accumulators: each CTA's TMEM holds its own 128
// queries x 256 rows, i.e. the epilogue is the single-CTA one.
// Protocol: only the leader (rank 0) issues MMAs.  Both producers signal the LEADER's `full`
// barrier (TMA .cta_group::2); the leader's commits arrive on `empty` / `tmem_full` in BOTH CTAs
// (multicast); both epilogues arrive on the LEADER's `tmem_empty` (one lane per warp).
constexpr int PAIR_BH_BYTES = (BN / 2) * BK * 2;        // 16 KiB: half pool tile per k-block
constexpr int PAIR_STAGE_BYTES = A_BYTES + PAIR_BH_BYTES;   // streaming layout
constexpr int PAIR_STAGES = 6;
constexpr int PAIR_DATA_BYTES = PAIR_STAGES * PAIR_STAGE_BYTES;   // 192 KiB (streaming)
constexpr int PAIR_RES_DATA_BYTES = RES_MAX_KB * A_BYTES + PAIR_STAGES * PAIR_BH_BYTES;  // 224 KiB
constexpr int PAIR_SMEM_BYTES = PAIR_RES_DATA_BYTES + 1024 + 256;
static_assert(PAIR_SMEM_BYTES <= 232448, "pair layout exceeds the 227 KiB CTA limit");
constexpr uint32_t kIdescPair = (1u << 4) | (static_cast<uint32_t>(BN >> 3) << 17) |
                                (static_cast<uint32_t>((2 * BM) >> 4) << 24);

__global__ void __launch_bounds__(NUM_THREADS, 1)
k_nns_coarse_pair(const __grid_constant__ CUtensorMap tmap_q,
                  const __grid_constant__ CUtensorMap tmap_p /* box rows BN / 2 */, CoarseParams prm) {
  const int crank = static_cast<int>(cluster_ctarank());
  const bool leader = crank == 0;
  const int cid = static_cast<int>(cluster_id_x());
  const int ncl = static_cast<int>(num_clusters_x());
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const bool a_res = prm.a_resident != 0;
  const uint32_t bar_base = smem_base + (a_res ? PAIR_RES_DATA_BYTES : PAIR_DATA_BYTES);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (PAIR_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * PAIR_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * PAIR_STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * PAIR_STAGES + 4);
  const uint32_t afull_bar = bar_base + 8u * (2 * PAIR_STAGES + 6);
  auto smem_a = [&](int s, int kb) {
    return a_res ? smem_base + kb * A_BYTES : smem_base + s * PAIR_STAGE_BYTES;
  };
  auto smem_b = [&](int s) {
    return a_res ? smem_base + RES_MAX_KB * A_BYTES + s * PAIR_BH_BYTES
                 : smem_base + s * PAIR_STAGE_BYTES + A_BYTES;
  };
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_p) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) {
      mbar_init(full_bar(s), 1);    // leader's copy is the one in use
      mbar_init(empty_bar(s), 1);   // leader's commit, multicast to both CTAs
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 16);   // 8 epilogue warps x 2 CTAs (leader's copy in use)
    }
    mbar_init(afull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(tmem_slot), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int num_kb = prm.num_kb;
  const int q_tile_row0 = prm.q_row0 + crank * BM;
  long long* dbg = (prm.dbg != nullptr && blockIdx.x == 0) ? prm.dbg : nullptr;
  long long w_acc0 = 0, w_acc1 = 0;
  auto timed_wait = [&](uint32_t bar, uint32_t parity, long long& acc) {
    if (dbg) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      PipeState ps;
      const uint32_t afull_leader = mapa_shared(afull_bar, 0);
      if (a_res) {
        if (leader) mbar_expect_tx(afull_bar, 2u * static_cast<uint32_t>(num_kb) * A_BYTES);
        for (int kb = 0; kb < num_kb; ++kb)
          tma_load_2d_2sm(smem_a(0, kb), &tmap_q, afull_leader, kb * BK, q_tile_row0, kEvictLast);
      }
      // the ring holds 96 KB of pool rows per SM - less than one HBM bandwidth-delay product at
      // 44 GB/s per SM - so this CTA's half tiles are pulled into L2 `pf` tiles ahead
      const int pf = prm.l2_prefetch;
      for (int t = 0; t < pf; ++t) {
        const int tl = cid + t * ncl;
        if (tl < prm.num_tiles) {
          const int r0 = static_cast<int>(static_cast<int64_t>(tl) * prm.tile_stride * BN);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_prefetch_l2_2d(&tmap_p, kb * BK, r0 + crank * (BN / 2));
        }
      }
      for (int tile = cid; tile < prm.num_tiles; tile += ncl) {
        const int row0 = static_cast<int>(static_cast<int64_t>(tile) * prm.tile_stride * BN);
        if (pf > 0) {
          const int tl = tile + pf * ncl;
          if (tl < prm.num_tiles) {
            const int r0 = static_cast<int>(static_cast<int64_t>(tl) * prm.tile_stride * BN);
            for (int kb = 0; kb < num_kb; ++kb)
              tma_prefetch_l2_2d(&tmap_p, kb * BK, r0 + crank * (BN / 2));
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          timed_wait(empty_bar(ps.stage), ps.phase ^ 1, w_acc0);
          const uint32_t full_leader = mapa_shared(full_bar(ps.stage), 0);
          if (leader)
            mbar_expect_tx(full_bar(ps.stage), 2u * (a_res ? PAIR_BH_BYTES : PAIR_STAGE_BYTES));
          if (!a_res)
            tma_load_2d_2sm(smem_a(ps.stage, kb), &tmap_q, full_leader, kb * BK, q_tile_row0,
                            kEvictLast);
          tma_load_2d_2sm(smem_b(ps.stage), &tmap_p, full_leader, kb * BK,
                          row0 + crank * (BN / 2), kEvictFirst);
          ps.advance(PAIR_STAGES);
        }
      }
      if (dbg) dbg[0] += w_acc0;
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      PipeState ps;
      int it = 0;
      if (a_res) {
        mbar_wait(afull_bar, 0);
        tc_fence_after();
      }
      for (int tile = cid; tile < prm.num_tiles; tile += ncl, ++it) {
        const int buf = it & 1;
        timed_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1, w_acc1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          timed_wait(full_bar(ps.stage), ps.phase, w_acc0);
          tc_fence_after();
          const uint32_t a0 = smem_a(ps.stage, kb);
          const uint32_t b0 = smem_b(ps.stage);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t adesc = make_sw128_desc(a0 + k * (UMMA_K * 2));
            const uint64_t bdesc = make_sw128_desc(b0 + k * (UMMA_K * 2));
            umma_f16_2sm(d_tmem, adesc, bdesc, kIdescPair, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm_mc(empty_bar(ps.stage), 0x3);   // frees the slot in both CTAs
          ps.advance(PAIR_STAGES);
        }
        umma_commit_2sm_mc(tfull_bar(buf), 0x3);          // accumulators complete in both CTAs
      }
      if (dbg) { dbg[1] += w_acc0; dbg[2] += w_acc1; }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quarter = ew & 3;
    const int half = ew >> 2;
    const int q = crank * BM + quarter * 32 + lane;
    const bool q_valid = q < prm.nq;
    const int qs = q_valid ? q : 0;
    const float tau = !q_valid ? INFINITY : (prm.tau != nullptr ? prm.tau[q] : -INFINITY);
    const size_t slots = cand_slots(prm.nsub);
    uint2* my_seg = prm.cand + static_cast<size_t>(qs) * slots +
                    static_cast<size_t>(cid * 2 + half) * kSegCap;
    uint2* my_ovf = prm.cand + static_cast<size_t>(qs) * slots +
                    static_cast<size_t>(prm.nsub) * kSegCap;
    unsigned int* my_cnt = prm.cnt + static_cast<size_t>(qs) * (prm.nsub + 1);
    uint32_t* my_smax = prm.smax + static_cast<size_t>(qs) * prm.smax_stride;
    unsigned int my_count = 0;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    int it = 0;
    for (int tile = cid; tile < prm.num_tiles; tile += ncl, ++it) {
      const int buf = it & 1;
      const int row0 = static_cast<int>(static_cast<int64_t>(tile) * prm.tile_stride * BN);
      timed_wait(tfull_bar(buf), (it >> 1) & 1, w_acc0);
      tc_fence_after();
      const long long te0 = dbg ? clock64() : 0;
      scan_accumulator_half(prm, lane_taddr + static_cast<uint32_t>(buf * BN), half, row0, tile, tau,
                            q_valid, my_seg, my_ovf, my_cnt, my_smax, my_count);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(buf), 0));
      if (dbg) w_acc1 += clock64() - te0;
    }
    if (dbg && threadIdx.x == 128) { dbg[3] += w_acc0; dbg[4] += w_acc1; dbg[5] += it; }
    if (q_valid && prm.mode == 0)
      my_cnt[cid * 2 + half] = min(my_count, static_cast<unsigned int>(kSegCap));
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(TMEM_COLS)
                 : "memory");
  }
}

}  // namespace

int make_fp16_rowmajor_tmap(void* out_map, const void* base, int64_t rows, int cols_pad,
                            int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    return CSLAM_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols_pad), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(cols_pad) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(out_map), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld cols=%d box_rows=%d",
              static_cast<int>(r), static_cast<long long>(rows), cols_pad, box_rows);
    return CSLAM_ERR_CUDA;
  }
  return CSLAM_OK;
}

int launch_coarse_tc(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int grid,
                     cudaStream_t stream) {
  CSLAM_CUDA(cudaFuncSetAttribute(k_nns_coarse_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SMEM_BYTES_RES));
  if (prm.num_tiles <= 0 || grid <= 0) return CSLAM_OK;
  if (prm.a_resident && (prm.num_kb > RES_MAX_KB || prm.cluster < 2)) {
    set_error("coarse_tc: A-resident mode needs a cluster and dim_pad <= %d", RES_MAX_KB * BK);
    return CSLAM_ERR_INVALID;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned int>(grid));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = prm.a_resident ? SMEM_BYTES_RES : SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned int>(prm.cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = prm.cluster > 1 ? 1 : 0;
  CSLAM_CUDA(cudaLaunchKernelEx(&cfg, k_nns_coarse_tc, *reinterpret_cast<const CUtensorMap*>(tmap_q),
                                *reinterpret_cast<const CUtensorMap*>(tmap_p), prm));
  count_launch();
  return CSLAM_OK;
}

// SM-pair kernel: clusters of 2, tmap_p with box rows kCoarseBN / 2; grid = CTAs (even)
int launch_coarse_pair(const void* tmap_q, const void* tmap_p, const CoarseParams& prm, int grid,
                       cudaStream_t stream) {
  CSLAM_CUDA(cudaFuncSetAttribute(k_nns_coarse_pair, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  PAIR_SMEM_BYTES));
  if (prm.num_tiles <= 0 || grid <= 0) return CSLAM_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned int>(grid));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = PAIR_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CSLAM_CUDA(cudaLaunchKernelEx(&cfg, k_nns_coarse_pair, *reinterpret_cast<const CUtensorMap*>(tmap_q),
                                *reinterpret_cast<const CUtensorMap*>(tmap_p), prm));
  count_launch();
  return CSLAM_OK;
}

int coarse_tc_max_clusters(int cluster, int* out) {
  CSLAM_CUDA(cudaFuncSetAttribute(k_nns_coarse_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  SMEM_BYTES_RES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned int>(cluster) * 64);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES_RES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = static_cast<unsigned int>(cluster);
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  CSLAM_CUDA(cudaOccupancyMaxActiveClusters(&n, k_nns_coarse_tc, &cfg));
  *out = n;
  return CSLAM_OK;
}

}  // namespace cslam
