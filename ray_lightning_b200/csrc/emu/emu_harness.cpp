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

#include "../b2d_kernels.cuh"
#ifdef B2D_EMU_WITH_PIPE
#include "../b2d_pipe.cuh"
#endif

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
#ifdef B2D_EMU_WITH_PIPE
    case 5: if constexpr (W == 2 || W == 4 || W == 8) k2p_two_shot_pipe_kernel<W, BF16, false>(P, pipe_k); break;
    case 6: if constexpr (W == 2 || W == 4 || W == 8) k2p_two_shot_pipe_kernel<W, BF16, true>(P, pipe_k); break;
#endif
    default: break;
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

// bufs[r]: rank r's fp32 bucket, reduced in place.  algo: 1 one-shot, 2 two-shot, 3 two-shot NVLS, 5/6 pipelined.
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
