// b2d_pipe.cuh — K2P: the two-shot allreduce as a role-decoupled chunk pipeline.
//
// STATUS: opt-in (`two_shot_pipe` = 5, `nvls_pipe` = 6), unreachable from AUTO.  Written after round 1's GPU
// budget was spent: its logic is validated on CPU threads against the oracle (tests/test_kernel_emulation.py,
// the very source below compiled for the host) and its protocol by exhaustive interleaving
// (tests/test_protocol_model.py::explore_pipelined); it compiles for sm_100a (128 registers, 4 named
// barriers) but has NOT run on hardware yet — tests/test_gpu_experimental.py is gated accordingly.
//
// Why: in K2 every phase already runs at its hardware limit (HBM 6.5 TB/s, NVLink ~790 GB/s) but the
// phases run one after the other and each of the two barriers stalls the whole block for 10-60 us.  Here the
// block's 16 warps are split into three roles that run CONCURRENTLY and are coupled only by monotone counters:
//
//     S (stage)   for c: cast+scale chunk c of EVERY slice into the own staging buffer; publish cntA = base+c+1
//     R (reduce)  for c: wait until cntA of ALL ranks >= base+c+1; reduce chunk c of MY slice from all ranks in
//                        rank order (or one multimem.ld_reduce), write it back in place (or multimem.st);
//                        publish cntB = base+c+1
//     G (gather)  for c: wait until cntB of ALL ranks >= base+c+1; read chunk c of every slice from its owner
//                        (NVLS: from the own arena) and write the fp32 bucket
//
// so HBM traffic (S, G's stores) overlaps NVLink traffic (R, G's loads) inside every SM, and a role that waits for
// a slower rank does not stop the other two.  The protocol (with the double-buffered slot and the kernel end as the
// join of the three roles) is checked exhaustively in tests/test_protocol_model.py::explore_pipelined.
// Ownership is per BLOCK: block b owns runs of kPipeRun consecutive packs, run q of block b = slice-relative packs
// [(q*grid + b)*kPipeRun, +kPipeRun), in every slice and every role; a chunk = K consecutive runs of the block.
// Arithmetic is the same Acc<> code as K1/K2: bit-identical results (P2P variant).
#pragma once

#include "b2d_kernels.cuh"

namespace b2d {

constexpr int kPipeRun = 128;                   // packs per run (2 KiB of wire)
constexpr int kPipeTS = 160, kPipeTR = 160, kPipeTG = 192;   // threads per role (sum = kThreads)
static_assert(kPipeTS + kPipeTR + kPipeTG == kThreads, "roles must fill the block");

__device__ __forceinline__ void named_barrier(int id, int nthreads) {
#ifdef B2D_EMU
  emu_named_barrier(id, nthreads);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

// spin until *ptr >= target (wrap-safe); trap with diagnostics after timeout_ns
__device__ __forceinline__ void spin_until_ge(const uint32_t* ptr, uint32_t target, unsigned long long timeout_ns,
                                              Diag* diag, int rank, int peer) {
  uint32_t got = ld_flag(ptr);
  if (static_cast<int32_t>(got - target) >= 0) return;
  const unsigned long long t0 = global_timer_ns();
  unsigned spins = 0;
  while (static_cast<int32_t>((got = ld_flag(ptr)) - target) < 0) {
    if ((++spins & 0xffu) == 0 && timeout_ns != 0 && global_timer_ns() - t0 > timeout_ns) {
      if (diag != nullptr) {
        diag->rank = rank; diag->block = blockIdx.x; diag->peer = peer; diag->expect = target; diag->got = got;
        diag->code = 1;
        __threadfence_system();
      }
      __trap();
    }
  }
}

template <int W, bool BF16, bool NVLS>
__global__ void __launch_bounds__(kThreads, 1) k2p_two_shot_pipe_kernel(const __grid_constant__ ArParams P, int K) {
  static_assert(W == 2 || W == 4 || W == 8, "K2P is instantiated for world 2, 4, 8");
  constexpr int EPP = BF16 ? 8 : 4;
  constexpr int LPP = BF16 ? 2 : 1;              // 16-byte gradient loads per pack in the stage role
  const size_t npacks = (P.n + EPP - 1) / EPP;
  const size_t slice = (npacks + W - 1) / W;
  const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  const size_t L = kPipeRun;
  const size_t total_runs = (slice + L - 1) / L;
  const size_t my_runs = total_runs > static_cast<size_t>(b) ? (total_runs - b + G - 1) / G : 0;
  const int C = static_cast<int>((my_runs + K - 1) / K);     // chunks of this block (same on every rank)
  const size_t KL = static_cast<size_t>(K) * L;              // packs per slice and chunk
  Signal* self = P.peers.signal[P.rank];
  const uint32_t base = self->pbase[b];
  uint4* my_stage = reinterpret_cast<uint4*>(P.peers.arena[P.rank] + P.stage_off);
  trace_stamp(P.trace, 0);

  // slice-relative pack of item i of chunk c, or `slice` (invalid) when past the block's runs / the slice
  auto item_jj = [&](int c, size_t i) -> size_t {
    const size_t q = static_cast<size_t>(c) * K + i / L;
    if (i >= KL || q >= my_runs) return slice;
    const size_t jj = (q * G + b) * L + i % L;
    return jj < slice ? jj : slice;
  };

  if (tid < kPipeTS) {
    // ---------------------------------------------------------------- role S: stage
    const int ts = tid;
    constexpr int UI = (16 / (W * LPP)) > 0 ? 16 / (W * LPP) : 1;   // items per batch: 16 loads in flight
    constexpr int B = UI * W;
    for (int c = 0; c < C; ++c) {
      for (size_t i0 = ts; i0 < KL; i0 += static_cast<size_t>(kPipeTS) * UI) {
        size_t p[B];
        bool ok[B];
#pragma unroll
        for (int u = 0; u < UI; ++u) {
          const size_t jj = item_jj(c, i0 + static_cast<size_t>(u) * kPipeTS);
#pragma unroll
          for (int s = 0; s < W; ++s) {
            p[u * W + s] = static_cast<size_t>(s) * slice + jj;
            ok[u * W + s] = jj < slice && p[u * W + s] < npacks;
          }
        }
        stage_batch<BF16, B>(P.grad, P.n, my_stage, p, ok, P.scale);
      }
      named_barrier(1, kPipeTS);                      // the whole role is done with chunk c
      if (ts < W) {
        __threadfence_system();                       // release, cumulative over the named barrier
        st_flag(&P.peers.signal[ts]->cntA[b][P.rank], base + static_cast<uint32_t>(c) + 1u);
      }
    }
  } else if (tid < kPipeTS + kPipeTR) {
    // ---------------------------------------------------------------- role R: reduce my slice
    const int tr = tid - kPipeTS;
    const size_t sbase = static_cast<size_t>(P.rank) * slice;
    for (int c = 0; c < C; ++c) {
      if (tr < W) {
        spin_until_ge(&self->cntA[b][tr], base + static_cast<uint32_t>(c) + 1u, P.timeout_ns, P.diag, P.rank, tr);
        __threadfence_system();                       // acquire
      }
      named_barrier(2, kPipeTR);
      if constexpr (NVLS) {
        constexpr int U = kMaxLoadsInFlight;
        const uint4* mc = reinterpret_cast<const uint4*>(P.peers.mc_arena + P.stage_off);
        for (size_t i0 = tr; i0 < KL; i0 += static_cast<size_t>(kPipeTR) * U) {
          uint4 red[U];
          size_t p[U];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t jj = item_jj(c, i0 + static_cast<size_t>(u) * kPipeTR);
            p[u] = sbase + jj;
            ok[u] = jj < slice && p[u] < npacks;
            if (ok[u]) red[u] = BF16 ? multimem_ld_reduce_bf16x8(mc + p[u]) : multimem_ld_reduce_f32x4(mc + p[u]);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (ok[u]) multimem_st_v4(const_cast<uint4*>(mc) + p[u], red[u]);
        }
      } else {
        constexpr int U = kMaxLoadsInFlight / W;
        for (size_t i0 = tr; i0 < KL; i0 += static_cast<size_t>(kPipeTR) * U) {
          uint4 in[U][W];
          size_t p[U];
          bool ok[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const size_t jj = item_jj(c, i0 + static_cast<size_t>(u) * kPipeTR);
            p[u] = sbase + jj;
            ok[u] = jj < slice && p[u] < npacks;
            if (ok[u]) {
#pragma unroll
              for (int r = 0; r < W; ++r)
                in[u][r] = ld_peer_v4(reinterpret_cast<const uint4*>(P.peers.arena[r] + P.stage_off) + p[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (ok[u]) {
              Acc<BF16> acc;
              acc.set(in[u][0]);
#pragma unroll
              for (int r = 1; r < W; ++r) acc.add(in[u][r]);
              st_v4(my_stage + p[u], acc.wire());
            }
          }
        }
      }
      named_barrier(2, kPipeTR);                      // every reduced pack of chunk c is written
      if (tr < W) {
        __threadfence_system();
        st_flag(&P.peers.signal[tr]->cntB[b][P.rank], base + static_cast<uint32_t>(c) + 1u);
      }
    }
  } else {
    // ---------------------------------------------------------------- role G: gather + fp32 write-back
    const int tg = tid - kPipeTS - kPipeTR;
    constexpr int U = kMaxLoadsInFlight / W;
    for (int c = 0; c < C; ++c) {
      if (tg < W) {
        spin_until_ge(&self->cntB[b][tg], base + static_cast<uint32_t>(c) + 1u, P.timeout_ns, P.diag, P.rank, tg);
        __threadfence_system();
      }
      named_barrier(3, kPipeTG);
      for (size_t i0 = tg; i0 < KL; i0 += static_cast<size_t>(kPipeTG) * U) {
        uint4 in[U * W];
        size_t p[U * W];
        bool ok[U * W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const size_t jj = item_jj(c, i0 + static_cast<size_t>(u) * kPipeTG);
#pragma unroll
          for (int s = 0; s < W; ++s) {
            const int i = u * W + s;
            p[i] = static_cast<size_t>(s) * slice + jj;
            ok[i] = jj < slice && p[i] < npacks;
            if (ok[i]) {
              const unsigned char* src = NVLS ? P.peers.arena[P.rank] : P.peers.arena[s];
              in[i] = ld_peer_v4(reinterpret_cast<const uint4*>(src + P.stage_off) + p[i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < U * W; ++i) {
          if (ok[i]) {
            uint4 raw[EPP / 4];
            from_wire<BF16>(in[i], raw);
            grad_store<EPP>(P.grad, P.n, p[i], raw);
          }
        }
      }
    }
  }
  __syncthreads();                                    // join of the three roles
  if (tid == 0) self->pbase[b] = base + static_cast<uint32_t>(C);
  trace_stamp(P.trace, 1);
}

}  // namespace b2d
