// emu_harness.cpp — runs the libb2d kernels (the very sources nvcc compiles) on CPU threads.
// TEST INFRASTRUCTURE ONLY: built by tests/test_kernel_emulation.py with g++ -std=c++20 -DB2D_EMU -pthread.
// See emu/cuda_emu.h for what is and is not modelled.
#ifndef B2D_EMU
#define B2D_EMU 1
#endif
#include <pthread.h>

#include <cmath>
#include <functional>
#include <vector>

#include <thread>

#include "../b2d_kernels.cuh"
#include "../b2d_staged.cuh"
#include "../b2d_owner.cuh"

thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local emu::Block* emu_block = nullptr;

using namespace b2d;

namespace {

struct Group {
  int world = 0;
  size_t arena_bytes = 0;
  std::vector<unsigned char*> arena;
  unsigned char* fake_mc = nullptr;
};

struct ThreadArg {
  std::function<void()>* body;
  emu::Block* block;
  unsigned tid, bid, nthreads, nblocks;
};

void* thread_main(void* p) {
  ThreadArg* a = static_cast<ThreadArg*>(p);
  threadIdx.x = a->tid; blockIdx.x = a->bid; blockDim.x = a->nthreads; gridDim.x = a->nblocks;
  emu_block = a->block;
  (*a->body)();
  return nullptr;
}

// launch `bodies[r]` as a grid x block kernel for every rank r, all ranks concurrently; join
int launch_all(int world, int grid, int block, std::vector<std::function<void()>>& bodies) {
  std::vector<std::unique_ptr<emu::Block>> blocks;
  std::vector<ThreadArg> args(static_cast<size_t>(world) * grid * block);
  std::vector<pthread_t> tids(args.size());
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 512 * 1024);
  size_t k = 0;
  for (int r = 0; r < world; ++r)
    for (int b = 0; b < grid; ++b) {
      blocks.emplace_back(new emu::Block(block));
      for (int t = 0; t < block; ++t, ++k) {
        args[k] = ThreadArg{&bodies[r], blocks.back().get(), static_cast<unsigned>(t), static_cast<unsigned>(b),
                            static_cast<unsigned>(block), static_cast<unsigned>(grid)};
        if (pthread_create(&tids[k], &attr, thread_main, &args[k]) != 0) return -1;
      }
    }
  for (size_t i = 0; i < k; ++i) pthread_join(tids[i], nullptr);
  pthread_attr_destroy(&attr);
  return 0;
}

Peers make_peers(const Group& g) {
  Peers p{};
  for (int r = 0; r < g.world; ++r) {
    p.arena[r] = g.arena[r];
    p.signal[r] = reinterpret_cast<Signal*>(g.arena[r]);
  }
  p.mc_arena = g.fake_mc;
  return p;
}

template <int W, bool BF16>
void run_ar(int algo, const ArParams& P, int pipe_k) {
  (void)pipe_k;
  switch (algo) {
    case 1: k1_one_shot_kernel<W, BF16>(P); break;
    case 2: k2_two_shot_kernel<W, BF16, false>(P); break;
    case 3: k2_two_shot_kernel<W, BF16, true>(P); break;
    default: break;
  }
}

// one kernel of one rank: grid x block threads, joined before returning (a stream runs these one after another)
int launch_one(int grid, int block, std::function<void()> body) {
  std::vector<std::function<void()>> bodies(1, std::move(body));
  return launch_all(1, grid, block, bodies);
}

template <int W>
void run_exch(const ExParams& P, bool bf16, bool nvls, bool inplace) {
  if (inplace) { if (nvls) exch_kernel<W, false, true, true>(P); else exch_kernel<W, false, false, true>(P); }
  else if (bf16) { if (nvls) exch_kernel<W, true, true, false>(P); else exch_kernel<W, true, false, false>(P); }
  else { if (nvls) exch_kernel<W, false, true, false>(P); else exch_kernel<W, false, false, false>(P); }
}
void run_exch_w(int world, bool generic, const ExParams& P, bool bf16, bool nvls, bool inplace) {
  if (generic) return run_exch<0>(P, bf16, nvls, inplace);
  switch (world) {
    case 2: return run_exch<2>(P, bf16, nvls, inplace);
    case 4: return run_exch<4>(P, bf16, nvls, inplace);
    case 8: return run_exch<8>(P, bf16, nvls, inplace);
    default: return run_exch<0>(P, bf16, nvls, inplace);
  }
}

}  // namespace

extern "C" {

void* emu_group_create(int world, size_t arena_bytes) {
  Group* g = new Group();
  g->world = world;
  g->arena_bytes = arena_bytes;
  for (int r = 0; r < world; ++r) {
    void* p = nullptr;
    if (posix_memalign(&p, 4096, arena_bytes) != 0) return nullptr;
    std::memset(p, 0, arena_bytes);
    g->arena.push_back(static_cast<unsigned char*>(p));
  }
  g->fake_mc = reinterpret_cast<unsigned char*>(static_cast<uintptr_t>(1) << 44);   // never dereferenced directly
  return g;
}

void emu_group_destroy(void* h) {
  Group* g = static_cast<Group*>(h);
  for (auto p : g->arena) free(p);
  delete g;
}

size_t emu_signal_bytes(void) { return kSignalBytes; }

// The staged exchange (b2d_staged.cuh) of one bucket, cut into chunks of `chunk_packs`, epochs epoch0.. (monotone
// over calls, like ctx->epoch in b2d.cu).  wire_off: byte offset of the staging region in every arena; with
// `inplace` the fp32 buckets themselves live there (bufs is ignored).  `order`:
//   0  every rank is one in-order stream S,X,W,U per chunk; the ranks run concurrently
//   1  fully serialised, phase-major (S of every rank, then X of every rank, then W+U): what a serialising
//      profiler makes of the single-GPU loopback ranks
//   2  three concurrent streams per rank (all S | all X | all W+U), coupled ONLY by the staged / published
//      flags — more freedom than the event-ordered streams of b2d.cu allow
int emu_staged_allreduce(void* h, int nvls, int bf16, int inplace, float** bufs, size_t n, float scale, size_t wire_off,
                         size_t chunk_packs, int st_grid, int ex_grid, unsigned epoch0, int order, int use_generic_w) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  const size_t epp = bf16 ? 8 : 4;
  const size_t npacks = (n + epp - 1) / epp;
  if (wire_off + npacks * 16 > g->arena_bytes) return -4;
  const int nchunks = static_cast<int>((npacks + chunk_packs - 1) / chunk_packs);
  emu::Multicast& mc = emu::multicast();
  mc.fake_base = g->fake_mc; mc.world = world;
  for (int r = 0; r < world; ++r) mc.arena[r] = g->arena[r];
  const Peers peers = make_peers(*g);
  auto S = [&](int r, int c) {
    StParams P{};
    P.scale = scale; P.rank = r; P.world = world; P.peers = peers;
    const size_t p0 = static_cast<size_t>(c) * chunk_packs, pc = npacks - p0 < chunk_packs ? npacks - p0 : chunk_packs;
    if (inplace) {
      if (c != 0) return 0;
      P.epoch = epoch0 + nchunks - 1;
      return launch_one(1, 32, [P] { arrive_kernel(P); });
    }
    P.grad = bufs[r] + p0 * epp;
    P.n = (p0 + pc) * epp <= n ? pc * epp : n - p0 * epp;
    P.wire = reinterpret_cast<uint4*>(g->arena[r] + wire_off) + p0;
    P.epoch = epoch0 + c;
    return bf16 ? launch_one(st_grid, kStThreads, [P] { stage_kernel<true>(P); }) : launch_one(st_grid, kStThreads, [P] { stage_kernel<false>(P); });
  };
  auto X = [&](int r, int c) {
    ExParams P{};
    P.scale = scale; P.rank = r; P.world = world; P.peers = peers; P.timeout_ns = 120ull * 1000000000ull; P.diag = nullptr;
    const size_t p0 = static_cast<size_t>(c) * chunk_packs, pc = npacks - p0 < chunk_packs ? npacks - p0 : chunk_packs;
    P.wire_off = wire_off + p0 * 16; P.npacks = pc; P.epoch = epoch0 + c;
    P.n_valid = inplace ? (n - p0 * 4 < pc * 4 ? n - p0 * 4 : pc * 4) : 0;
    return launch_one(ex_grid, kExThreads, [=] { run_exch_w(world, use_generic_w != 0, P, bf16 != 0, nvls != 0, inplace != 0); });
  };
  auto WU = [&](int r, int c) {
    ExParams P{};
    P.rank = r; P.world = world; P.peers = peers; P.timeout_ns = 120ull * 1000000000ull; P.epoch = epoch0 + c;
    int rc = launch_one(1, 32, [P] { wait_published_kernel(P); });
    if (rc != 0 || inplace) return rc;
    StParams Q{};
    Q.scale = scale; Q.rank = r; Q.world = world; Q.peers = peers;
    const size_t p0 = static_cast<size_t>(c) * chunk_packs, pc = npacks - p0 < chunk_packs ? npacks - p0 : chunk_packs;
    Q.grad = bufs[r] + p0 * epp;
    Q.n = (p0 + pc) * epp <= n ? pc * epp : n - p0 * epp;
    Q.wire = reinterpret_cast<uint4*>(g->arena[r] + wire_off) + p0;
    Q.epoch = epoch0 + c;
    return bf16 ? launch_one(st_grid, kStThreads, [Q] { unstage_kernel<true>(Q); }) : launch_one(st_grid, kStThreads, [Q] { unstage_kernel<false>(Q); });
  };
  std::atomic<int> bad{0};
  if (order == 1) {
    for (int c = 0; c < nchunks; ++c) for (int r = 0; r < world; ++r) if (S(r, c) != 0) return -1;
    for (int c = 0; c < nchunks; ++c) for (int r = 0; r < world; ++r) if (X(r, c) != 0) return -1;
    for (int c = 0; c < nchunks; ++c) for (int r = 0; r < world; ++r) if (WU(r, c) != 0) return -1;
    return 0;
  }
  std::vector<std::thread> streams;
  for (int r = 0; r < world; ++r) {
    if (order == 0) {
      streams.emplace_back([&, r] { for (int c = 0; c < nchunks; ++c) if (S(r, c) != 0 || X(r, c) != 0 || WU(r, c) != 0) bad = 1; });
    } else {
      streams.emplace_back([&, r] { for (int c = 0; c < nchunks; ++c) if (S(r, c) != 0) bad = 1; });
      streams.emplace_back([&, r] { for (int c = 0; c < nchunks; ++c) if (X(r, c) != 0) bad = 1; });
      streams.emplace_back([&, r] { for (int c = 0; c < nchunks; ++c) if (WU(r, c) != 0) bad = 1; });
    }
  }
  for (auto& t : streams) t.join();
  return bad.load() ? -1 : 0;
}

// K11 + K12 (b2d_owner.cuh): one reduce bucket given as an owner-sorted segment table (seg_start in packs of the wire
// format, cumulative, nseg + 1 entries; owner_pack[world + 1]).  order as in emu_staged_allreduce (0 / 1).
int emu_reduce_to_owner(void* h, int bf16, int nvls, float** grads, float** reduced, const long long* shard_off,
                        const long long* seg_flat_off, const unsigned* seg_start, int nseg, const unsigned* owner_pack,
                        size_t wire_off, float scale, int zero_grads, int accumulate, unsigned epoch, int order, int use_generic_w) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  if (wire_off + static_cast<size_t>(seg_start[nseg]) * 16 > g->arena_bytes) return -4;
  emu::Multicast& mc = emu::multicast();
  mc.fake_base = g->fake_mc; mc.world = world;
  for (int r = 0; r < world; ++r) mc.arena[r] = g->arena[r];
  const Peers peers = make_peers(*g);
  auto mk = [&](int r) {
    SegParams P{};
    P.seg_flat_off = seg_flat_off; P.seg_start = seg_start; P.nseg = nseg;
    for (int i = 0; i <= B2D_MAX_WORLD; ++i) P.owner_pack[i] = owner_pack[i <= world ? i : world];
    P.grads = grads[r]; P.reduced = reduced[r]; P.shard_lo = shard_off[r]; P.wire_off = wire_off; P.scale = scale;
    P.zero_grads = zero_grads; P.accumulate = accumulate; P.rank = r; P.world = world; P.epoch = epoch;
    P.timeout_ns = 120ull * 1000000000ull; P.diag = nullptr; P.peers = peers;
    return P;
  };
  auto S = [&](int r) {
    const SegParams P = mk(r);
    return bf16 ? launch_one(2, kStThreads, [P] { seg_stage_kernel<true>(P); }) : launch_one(2, kStThreads, [P] { seg_stage_kernel<false>(P); });
  };
  auto X = [&](int r) {
    const SegParams P = mk(r);
    auto run = [=] {
      auto go = [&](auto w) {
        constexpr int W = decltype(w)::value;
        if (bf16) { if (nvls) seg_reduce_kernel<W, true, true>(P); else seg_reduce_kernel<W, true, false>(P); }
        else { if (nvls) seg_reduce_kernel<W, false, true>(P); else seg_reduce_kernel<W, false, false>(P); }
      };
      if (use_generic_w) go(std::integral_constant<int, 0>{});
      else if (world == 2) go(std::integral_constant<int, 2>{});
      else if (world == 4) go(std::integral_constant<int, 4>{});
      else if (world == 8) go(std::integral_constant<int, 8>{});
      else go(std::integral_constant<int, 0>{});
    };
    return launch_one(2, kExThreads, run);
  };
  if (order == 1) {
    for (int r = 0; r < world; ++r) if (S(r) != 0) return -1;
    for (int r = 0; r < world; ++r) if (X(r) != 0) return -1;
    return 0;
  }
  std::atomic<int> bad{0};
  std::vector<std::thread> streams;
  for (int r = 0; r < world; ++r) streams.emplace_back([&, r] { if (S(r) != 0 || X(r) != 0) bad = 1; });
  for (auto& t : streams) t.join();
  return bad.load() ? -1 : 0;
}

// K13: params live in every arena at param_off; m, v, reduced are the ranks' own-shard buffers.  One Adam group per
// rank covering [glo[r], ghi[r]) of its shard (ngroups = 0: push only).
int emu_adam_push(void* h, int nvls, size_t param_off, float** m, float** v, float** reduced, size_t n, const long long* shard_off,
                  int ngroups, const long long* glo, const long long* ghi, float lr, float beta1, float beta2, float eps, float wd,
                  int step, int adamw, unsigned epoch, int order, int use_generic_w) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  if (param_off + n * 4 > g->arena_bytes) return -4;
  emu::Multicast& mc = emu::multicast();
  mc.fake_base = g->fake_mc; mc.world = world;
  for (int r = 0; r < world; ++r) mc.arena[r] = g->arena[r];
  const Peers peers = make_peers(*g);
  auto X = [&](int r) {
    PushParams P{};
    P.params = reinterpret_cast<float*>(g->arena[r] + param_off); P.param_off = param_off;
    P.exp_avg = m ? m[r] : nullptr; P.exp_avg_sq = v ? v[r] : nullptr; P.reduced = reduced ? reduced[r] : nullptr;
    P.lo = shard_off[r]; P.hi = shard_off[r + 1]; P.ngroups = ngroups;
    if (ngroups > 0) {
      P.group_lo[0] = glo[r]; P.group_hi[0] = ghi[r];
      AdamConsts& a = P.group[0];
      a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = wd;
      a.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
      a.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
      double b1p = 1.0, b2p = 1.0;
      for (int i = 0; i < step; ++i) { b1p *= static_cast<double>(beta1); b2p *= static_cast<double>(beta2); }
      a.step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - b1p));
      a.inv_bc2_sqrt = 1.0f / static_cast<float>(std::sqrt(1.0 - b2p));
      a.decay_mul = static_cast<float>(1.0 - static_cast<double>(lr) * static_cast<double>(wd));
      a.adamw = adamw;
    }
    P.rank = r; P.world = world; P.epoch = epoch; P.peers = peers;
    auto run = [=] {
      if (use_generic_w) { if (nvls) adam_push_kernel<0, true>(P); else adam_push_kernel<0, false>(P); }
      else if (world == 2) { if (nvls) adam_push_kernel<2, true>(P); else adam_push_kernel<2, false>(P); }
      else if (world == 4) { if (nvls) adam_push_kernel<4, true>(P); else adam_push_kernel<4, false>(P); }
      else { if (nvls) adam_push_kernel<0, true>(P); else adam_push_kernel<0, false>(P); }
    };
    return launch_one(2, kExThreads, run);
  };
  auto Wt = [&](int r) {
    ExParams P{};
    P.rank = r; P.world = world; P.peers = peers; P.timeout_ns = 120ull * 1000000000ull; P.epoch = epoch;
    return launch_one(1, 32, [P] { wait_published_kernel(P); });
  };
  if (order == 1) {
    for (int r = 0; r < world; ++r) if (X(r) != 0) return -1;
    for (int r = 0; r < world; ++r) if (Wt(r) != 0) return -1;
    return 0;
  }
  std::atomic<int> bad{0};
  std::vector<std::thread> streams;
  for (int r = 0; r < world; ++r) streams.emplace_back([&, r] { if (X(r) != 0 || Wt(r) != 0) bad = 1; });
  for (auto& t : streams) t.join();
  return bad.load() ? -1 : 0;
}

// K14: one optimizer step of a bucket whose parameters (and their state tensors) are separate allocations.
int emu_bucket_optim(float** params, float** state1, float** state2, const unsigned* seg_start, int nseg, const float* grads,
                     size_t n, int kind, float lr, float momentum, float wd, float beta1, float beta2, float eps, int step, int adamw) {
  OptimParams P{};
  P.param_ptr = params; P.state1_ptr = state1; P.state2_ptr = state2; P.seg_start = seg_start; P.nseg = nseg;
  P.grads = grads; P.n = n; P.kind = kind; P.lr = lr; P.momentum = momentum; P.weight_decay = wd;
  if (kind == 1) {
    AdamConsts& a = P.adam;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = wd;
    a.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
    a.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
    double b1p = 1.0, b2p = 1.0;
    for (int i = 0; i < step; ++i) { b1p *= static_cast<double>(beta1); b2p *= static_cast<double>(beta2); }
    a.step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - b1p));
    a.inv_bc2_sqrt = 1.0f / static_cast<float>(std::sqrt(1.0 - b2p));
    a.decay_mul = static_cast<float>(1.0 - static_cast<double>(lr) * static_cast<double>(wd));
    a.adamw = adamw;
  }
  return launch_one(2, kStThreads, [P] { bucket_optim_kernel(P); });
}

// bufs[r]: rank r's fp32 bucket, reduced in place.  algo: 1 one-shot, 2 two-shot, 3 two-shot NVLS (fused).
// parity selects the half of the (single) double-buffered slot, exactly like get_slot() in b2d.cu.
int emu_allreduce(void* h, int algo, int bf16, float** bufs, size_t n, float scale, int grid, int parity, int use_generic_w,
                  int pipe_k) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  const size_t epp = bf16 ? 8 : 4;
  const size_t npacks = (n + epp - 1) / epp, slice = (npacks + world - 1) / world;
  const size_t half = (slice * world * 16 + 255) / 256 * 256;
  if (kSignalBytes + 2 * half > g->arena_bytes) return -4;
  emu::Multicast& mc = emu::multicast();
  mc.fake_base = g->fake_mc; mc.world = world;
  for (int r = 0; r < world; ++r) mc.arena[r] = g->arena[r];
  std::vector<ArParams> params(world);
  std::vector<std::function<void()>> bodies(world);
  for (int r = 0; r < world; ++r) {
    ArParams P{};
    P.trace = nullptr; P.grad = bufs[r]; P.n = n; P.stage_off = kSignalBytes + (parity & 1) * half; P.scale = scale;
    P.rank = r; P.world = world; P.timeout_ns = 60ull * 1000000000ull; P.diag = nullptr; P.peers = make_peers(*g);
    params[r] = P;
    const ArParams* pp = &params[r];
    if (use_generic_w) {
      bodies[r] = bf16 ? std::function<void()>([=] { run_ar<0, true>(algo, *pp, pipe_k); })
                       : std::function<void()>([=] { run_ar<0, false>(algo, *pp, pipe_k); });
    } else {
      switch (world) {
        case 2: bodies[r] = bf16 ? std::function<void()>([=] { run_ar<2, true>(algo, *pp, pipe_k); })
                                 : std::function<void()>([=] { run_ar<2, false>(algo, *pp, pipe_k); }); break;
        case 4: bodies[r] = bf16 ? std::function<void()>([=] { run_ar<4, true>(algo, *pp, pipe_k); })
                                 : std::function<void()>([=] { run_ar<4, false>(algo, *pp, pipe_k); }); break;
        case 8: bodies[r] = bf16 ? std::function<void()>([=] { run_ar<8, true>(algo, *pp, pipe_k); })
                                 : std::function<void()>([=] { run_ar<8, false>(algo, *pp, pipe_k); }); break;
        default: return -1;
      }
    }
  }
  return launch_all(world, grid, kThreads, bodies);
}

// K0: world 1, no peers
int emu_k0(float* buf, size_t n, float scale, int bf16, int grid) {
  std::vector<std::function<void()>> bodies(1);
  bodies[0] = bf16 ? std::function<void()>([=] { k0_cast_scale_kernel<true>(buf, n, scale); })
                   : std::function<void()>([=] { k0_cast_scale_kernel<false>(buf, n, scale); });
  return launch_all(1, grid, kThreads, bodies);
}

// Fused sharded step.  params live in every arena at `param_off`; grads[r], m[r], v[r] are plain buffers.
int emu_sharded_step(void* h, int bf16, float** grads, size_t param_off, float** m, float** v, size_t n,
                     const long long* shard_off, float scale, float lr, float beta1, float beta2, float eps, float wd,
                     int step, int adamw, int zero_grads, int grid, int parity, int use_generic_w) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  const size_t half = (n * (bf16 ? 2 : 4) + 255) / 256 * 256;
  const size_t stage_base = param_off + ((n * 4 + 255) / 256 * 256);
  if (stage_base + 2 * half > g->arena_bytes) return -4;
  std::vector<ShParams> params(world);
  std::vector<std::function<void()>> bodies(world);
  for (int r = 0; r < world; ++r) {
    ShParams P{};
    P.grads = grads[r]; P.grads_rw = zero_grads ? grads[r] : nullptr;
    P.params = reinterpret_cast<float*>(g->arena[r] + param_off); P.param_off = param_off;
    P.exp_avg = m[r]; P.exp_avg_sq = v[r]; P.rs_out = nullptr; P.n = n;
    for (int i = 0; i <= world; ++i) P.off[i] = shard_off[i];
    for (int i = world + 1; i <= B2D_MAX_WORLD; ++i) P.off[i] = shard_off[world];
    P.stage_off = stage_base + (parity & 1) * half; P.scale = scale; P.rank = r; P.world = world;
    P.do_stage_reduce = 1; P.do_adam = 1; P.do_gather = 1; P.end_barrier = 0;
    AdamConsts& a = P.adam;   // same host arithmetic as sharded_common() in b2d.cu
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = wd;
    a.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
    a.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
    double b1p = 1.0, b2p = 1.0;
    for (int i = 0; i < step; ++i) { b1p *= static_cast<double>(beta1); b2p *= static_cast<double>(beta2); }
    a.step_size = static_cast<float>(static_cast<double>(lr) / (1.0 - b1p));
    a.inv_bc2_sqrt = 1.0f / static_cast<float>(std::sqrt(1.0 - b2p));
    a.decay_mul = static_cast<float>(1.0 - static_cast<double>(lr) * static_cast<double>(wd));
    a.adamw = adamw;
    P.timeout_ns = 60ull * 1000000000ull; P.diag = nullptr; P.peers = make_peers(*g);
    params[r] = P;
    const ShParams* pp = &params[r];
    auto mk = [&](auto fn) { return std::function<void()>([=] { fn(*pp); }); };
    if (use_generic_w) bodies[r] = bf16 ? mk(k456_sharded_kernel<0, true>) : mk(k456_sharded_kernel<0, false>);
    else if (world == 2) bodies[r] = bf16 ? mk(k456_sharded_kernel<2, true>) : mk(k456_sharded_kernel<2, false>);
    else if (world == 4) bodies[r] = bf16 ? mk(k456_sharded_kernel<4, true>) : mk(k456_sharded_kernel<4, false>);
    else if (world == 8) bodies[r] = bf16 ? mk(k456_sharded_kernel<8, true>) : mk(k456_sharded_kernel<8, false>);
    else return -1;
  }
  return launch_all(world, grid, kThreads, bodies);
}

// K4 alone (reduce-scatter to owner, fp32 out) and K6 alone (all-gather of a flat arena buffer, with its end barrier)
int emu_reduce_scatter(void* h, int bf16, float** grads, float** outs, size_t n, const long long* shard_off, float scale,
                       size_t stage_base, int grid, int parity) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  const size_t half = (n * (bf16 ? 2 : 4) + 255) / 256 * 256;
  if (stage_base + 2 * half > g->arena_bytes) return -4;
  std::vector<ShParams> params(world);
  std::vector<std::function<void()>> bodies(world);
  for (int r = 0; r < world; ++r) {
    ShParams P{};
    P.grads = grads[r]; P.rs_out = outs[r]; P.n = n;
    for (int i = 0; i <= world; ++i) P.off[i] = shard_off[i];
    for (int i = world + 1; i <= B2D_MAX_WORLD; ++i) P.off[i] = shard_off[world];
    P.stage_off = stage_base + (parity & 1) * half; P.scale = scale; P.rank = r; P.world = world;
    P.do_stage_reduce = 1; P.do_adam = 0; P.do_gather = 0;
    P.timeout_ns = 60ull * 1000000000ull; P.peers = make_peers(*g);
    params[r] = P;
    const ShParams* pp = &params[r];
    bodies[r] = bf16 ? std::function<void()>([=] { k456_sharded_kernel<0, true>(*pp); })
                     : std::function<void()>([=] { k456_sharded_kernel<0, false>(*pp); });
  }
  return launch_all(world, grid, kThreads, bodies);
}

int emu_allgather(void* h, size_t buf_off, size_t n, const long long* shard_off, int grid) {
  Group* g = static_cast<Group*>(h);
  const int world = g->world;
  std::vector<ShParams> params(world);
  std::vector<std::function<void()>> bodies(world);
  for (int r = 0; r < world; ++r) {
    ShParams P{};
    P.params = reinterpret_cast<float*>(g->arena[r] + buf_off); P.param_off = buf_off; P.n = n;
    for (int i = 0; i <= world; ++i) P.off[i] = shard_off[i];
    for (int i = world + 1; i <= B2D_MAX_WORLD; ++i) P.off[i] = shard_off[world];
    P.rank = r; P.world = world; P.do_gather = 1; P.end_barrier = 1;
    P.timeout_ns = 60ull * 1000000000ull; P.peers = make_peers(*g);
    params[r] = P;
    const ShParams* pp = &params[r];
    bodies[r] = std::function<void()>([=] { k456_sharded_kernel<0, false>(*pp); });
  }
  return launch_all(world, grid, kThreads, bodies);
}

float* emu_arena_ptr(void* h, int rank, size_t off) {
  return reinterpret_cast<float*>(static_cast<Group*>(h)->arena[rank] + off);
}

}  // extern "C"
