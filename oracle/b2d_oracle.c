/*
 * b2d_oracle.c — plain-C restatement of the per-bucket arithmetic of the reference's gradient
 * sync.  TEST INFRASTRUCTURE ONLY: linked by nothing under ray_lightning_b200/; loaded (ctypes)
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
 *
 * Follows (see oracle/ddp_oracle.py for the full citations):
 *   [R1] torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:18-54   divide, fp32 SUM
 *   [R2] torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:57-93,116-134
 *        buffer.to(bf16).div_(W) -> SUM -> copy back to fp32
 *   [R3] torch/nn/parallel/distributed.py:1183-1281  bucket assignment by size (+ reversal)
 *   [R4] FairScale OSS.partition_parameters (recalled) / zero_redundancy_optimizer.py:651-722
 *   [R5] torch/optim/adam.py:530-547
 * reached from ray_lightning/ray_ddp.py:112-116 (**ddp_kwargs -> DistributedDataParallel) and
 * ray_lightning/ray_ddp_sharded.py:12-13.
 *
 * Parity pinning: checked against fixtures produced by the reference's real implementation
 * (torch DDP over gloo) in tests/golden/, see tests/test_oracle.py.
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (no FMA contraction: every add/mul rounds).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* fp32 -> bf16 -> fp32, round to nearest even; NaN stays NaN (quiet), like c10::BFloat16 */
float oracle_bf16_round(float x) {
  uint32_t u = f2u(x);
  if ((u & 0x7fffffffu) > 0x7f800000u) return u2f((u | 0x00400000u) & 0xffff0000u);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return u2f(u & 0xffff0000u);
}

/* [R2] value a rank puts on the wire: bf16(fp32(bf16(g)) * scale), scale = 1.0f / W */
float oracle_wire_bf16(float g, float scale) { return oracle_bf16_round(oracle_bf16_round(g) * scale); }

/* contract of B2D_WIRE_BF16: out = fp32(bf16(sum_{r in order} wire(g_r))), fp32 accumulate */
void oracle_allreduce_bf16(const float* const* per_rank, int world, size_t n, float scale, float* out) {
  for (size_t i = 0; i < n; ++i) {
    float acc = oracle_wire_bf16(per_rank[0][i], scale);
    for (int r = 1; r < world; ++r) acc = acc + oracle_wire_bf16(per_rank[r][i], scale);
    out[i] = oracle_bf16_round(acc);
  }
}

/* contract of B2D_WIRE_FP32 == [R1]: out = sum_{r in order} g_r * scale */
void oracle_allreduce_fp32(const float* const* per_rank, int world, size_t n, float scale, float* out) {
  for (size_t i = 0; i < n; ++i) {
    float acc = per_rank[0][i] * scale;
    for (int r = 1; r < world; ++r) acc = acc + per_rank[r][i] * scale;
    out[i] = acc;
  }
}

/* [R5] one Adam step on a flat fp32 tensor, python-float (double) scalar arithmetic */
void oracle_adam(float* p, const float* g_in, float* m, float* v, size_t n, int step, double lr,
                 double beta1, double beta2, double eps, double weight_decay, int adamw) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  const float w1 = (float)(1.0 - beta1), w2 = (float)(1.0 - beta2), b2 = (float)beta2;
  for (size_t i = 0; i < n; ++i) {
    float g = g_in[i];
    if (adamw) p[i] = p[i] * (float)(1.0 - lr * weight_decay);
    else if (weight_decay != 0.0) g = g + (float)weight_decay * p[i];
    m[i] = m[i] + w1 * (g - m[i]);
    v[i] = v[i] * b2 + w2 * g * g;
    const float denom = sqrtf(v[i]) / bc2_sqrt + (float)eps;
    p[i] = p[i] + (-step_size) * (m[i] / denom);
  }
}

/* [R3] greedy in-order bucket assignment, one dtype/device; limits[] in bytes, the last limit
 * repeats.  bucket_of[i] = bucket index BEFORE reversal; returns the bucket count. */
int oracle_bucket_assignment(const int64_t* numels, int nparams, int elem_size, const int64_t* limits,
                             int nlimits, int* bucket_of) {
  int nb = 0, li = 0, open = 0;
  int64_t cur = 0;
  for (int i = 0; i < nparams; ++i) {
    bucket_of[i] = nb;
    open = 1;
    cur += numels[i] * elem_size;
    if (cur >= limits[li < nlimits ? li : nlimits - 1]) { nb++; cur = 0; open = 0; li++; }
  }
  return nb + open;
}

/* [R4] FairScale rule: declaration order, smallest running size, first minimum */
void oracle_partition_fairscale(const int64_t* numels, int nparams, int world, int* owner) {
  int64_t sizes[64] = {0};
  for (int i = 0; i < nparams; ++i) {
    int best = 0;
    for (int r = 1; r < world; ++r) if (sizes[r] < sizes[best]) best = r;
    owner[i] = best;
    sizes[best] += numels[i];
  }
}
