// SPDX-License-Identifier: Apache-2.0
//
// filter_kernels.cu -- device interpreter for the post-merge predicate (K5 in SURVEY.md):
// one thread per merged row runs a tiny postfix program with SQL three-valued logic and writes
// keep[u] = (predicate is TRUE).
#include <cuda_runtime.h>

#include "filter_prog.h"

namespace lsb {

constexpr uint32_t kNoRowF = 0xffffffffu;
enum : uint8_t { TV_FALSE = 0, TV_TRUE = 1, TV_NULL = 2 };

__device__ __forceinline__ int cmp_to_bool(int c, int op) {
  switch (op) {
    case 0: return c == 0;
    case 1: return c != 0;
    case 2: return c > 0;
    case 3: return c >= 0;
    case 4: return c < 0;
    default: return c <= 0;
  }
}

__device__ __forceinline__ int64_t total_f64(double d) {
  int64_t b = __double_as_longlong(d);
  return b ^ (int64_t)(((uint64_t)(b >> 63)) >> 1);
}
__device__ __forceinline__ int32_t total_f32(float f) {
  int32_t b = __float_as_int(f);
  return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
}

template <typename Q>
__device__ __forceinline__ int upper_idx64(const int64_t* arr, int n, Q q) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (arr[mid] <= q) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) filter_kernel(const FilterProgram* __restrict__ progp,
                                                     const uint32_t* __restrict__ plan_win,
                                                     const int64_t* __restrict__ out_base, int nb, int64_t U,
                                                     uint8_t* __restrict__ keep) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const FilterProgram& prog = *progp;
  uint8_t stack[kMaxFilterInstr];
  int sp = 0;
  int bucket = -1;
  for (int pc = 0; pc < prog.n_instr; ++pc) {
    const FilterInstr& in = prog.instr[pc];
    switch (in.opcode) {
      case FO_TRUE: stack[sp++] = TV_TRUE; break;
      case FO_FALSE: stack[sp++] = TV_FALSE; break;
      case FO_NULL: stack[sp++] = TV_NULL; break;
      case FO_NOT: {
        const uint8_t a = stack[sp - 1];
        stack[sp - 1] = a == TV_NULL ? TV_NULL : (a == TV_TRUE ? TV_FALSE : TV_TRUE);
      } break;
      case FO_AND: {
        const uint8_t b = stack[--sp], a = stack[sp - 1];
        stack[sp - 1] = (a == TV_FALSE || b == TV_FALSE) ? TV_FALSE
                        : (a == TV_NULL || b == TV_NULL) ? TV_NULL : TV_TRUE;
      } break;
      case FO_OR: {
        const uint8_t b = stack[--sp], a = stack[sp - 1];
        stack[sp - 1] = (a == TV_TRUE || b == TV_TRUE) ? TV_TRUE
                        : (a == TV_NULL || b == TV_NULL) ? TV_NULL : TV_FALSE;
      } break;
      default: {
        const FilterColDev& c = prog.cols[in.col];
        const uint32_t g = (c.win ? c.win : plan_win)[u];
        const bool is_null = g == kNoRowF || (c.validity && !((c.validity[g >> 3] >> (g & 7)) & 1));
        if (in.opcode == FO_ISNULL) { stack[sp++] = is_null ? TV_TRUE : TV_FALSE; break; }
        if (in.opcode == FO_ISNOTNULL) { stack[sp++] = is_null ? TV_FALSE : TV_TRUE; break; }
        if (is_null) { stack[sp++] = TV_NULL; break; }
        int cmp = 0;
        switch (c.cls) {
          case FC_INT: {
            int64_t v;
            switch (c.width) {
              case 1: v = reinterpret_cast<const int8_t*>(c.values)[g]; break;
              case 2: v = reinterpret_cast<const int16_t*>(c.values)[g]; break;
              case 4: v = reinterpret_cast<const int32_t*>(c.values)[g]; break;
              default: v = reinterpret_cast<const int64_t*>(c.values)[g]; break;
            }
            if (c.mul != 1) v *= c.mul;
            cmp = v < in.lit_i ? -1 : (v > in.lit_i ? 1 : 0);
          } break;
          case FC_UINT: {
            uint64_t v;
            switch (c.width) {
              case 1: v = reinterpret_cast<const uint8_t*>(c.values)[g]; break;
              case 2: v = reinterpret_cast<const uint16_t*>(c.values)[g]; break;
              case 4: v = reinterpret_cast<const uint32_t*>(c.values)[g]; break;
              default: v = reinterpret_cast<const uint64_t*>(c.values)[g]; break;
            }
            cmp = v < in.lit_u ? -1 : (v > in.lit_u ? 1 : 0);
          } break;
          case FC_F32: {
            const int32_t a = total_f32(reinterpret_cast<const float*>(c.values)[g]);
            const int32_t b = total_f32((float)in.lit_f);
            cmp = a < b ? -1 : (a > b ? 1 : 0);
          } break;
          case FC_F64: {
            const int64_t a = total_f64(reinterpret_cast<const double*>(c.values)[g]);
            const int64_t b = total_f64(in.lit_f);
            cmp = a < b ? -1 : (a > b ? 1 : 0);
          } break;
          case FC_BOOL: {
            const int v = (reinterpret_cast<const uint8_t*>(c.values)[g >> 3] >> (g & 7)) & 1;
            cmp = v - (int)in.lit_i;
          } break;
          case FC_DEC128: {
            const uint64_t lo = reinterpret_cast<const uint64_t*>(c.values)[2 * (int64_t)g];
            const int64_t hi = reinterpret_cast<const int64_t*>(c.values)[2 * (int64_t)g + 1];
            const int64_t lhi = (int64_t)in.lit_hi;
            if (hi != lhi) cmp = hi < lhi ? -1 : 1;
            else cmp = lo < in.lit_lo ? -1 : (lo > in.lit_lo ? 1 : 0);
          } break;
          default: {  // FC_BYTES
            if (bucket < 0) bucket = upper_idx64(out_base, nb + 1, u);
            const int32_t* offs = reinterpret_cast<const int32_t*>(c.values);
            const int32_t o0 = offs[g], o1 = offs[(int64_t)g + 1];
            const uint8_t* p = c.data + c.bucket_data[bucket] + o0;
            const uint8_t* q = prog.lit_pool + in.lit_ofs;
            const int la = o1 - o0, lb = in.lit_len;
            const int n = la < lb ? la : lb;
            cmp = 0;
            for (int i = 0; i < n && cmp == 0; ++i) cmp = (int)p[i] - (int)q[i];
            if (cmp == 0) cmp = la - lb;
          } break;
        }
        stack[sp++] = cmp_to_bool(cmp, in.cmp) ? TV_TRUE : TV_FALSE;
      } break;
    }
  }
  keep[u] = (sp > 0 && stack[sp - 1] == TV_TRUE) ? 1 : 0;
}

void launch_filter_kernel(const FilterProgram* d_prog, const uint32_t* plan_win, const int64_t* out_base,
                          int nb, int64_t U, uint8_t* keep, cudaStream_t stream) {
  if (U <= 0) return;
  filter_kernel<<<(unsigned)((U + 255) / 256), 256, 0, stream>>>(d_prog, plan_win, out_base, nb, U, keep);
}

}  // namespace lsb
