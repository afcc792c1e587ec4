/*
 * xbit_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity checker, never the product).
 *
 * Plain-C CPU restatement of the reference's hot path (wejoncy/XbitOps): group-wise 2..8-bit
 * dequantisation to fp16 and the A16Wx GEMV "truth".  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's library.  The product
 * (xbitops_b200/csrc) never links, imports or falls back to it.
 *
 * Parity pin status: PINNED BY EXECUTING THE REFERENCE (the reference ships no golden vectors,
 * SURVEY.md 8(c)):
 *   - integer-exact and <=1-ulp fp16 against the reference's own CPU simulator
 *     (src/cpp_simulate.cc:568-691, compiled unmodified into oracle/_ref/ by oracle/Makefile),
 *     checked in tests/test_oracle.py and frozen in tests/golden/ (generator: tests/golden/make_golden.py);
 *   - bit-exact fp16 against the reference's GPU kernels built for compute_100
 *     (oracle/build_ref_gpu.sh) in tests/test_ref_gpu_parity.py (gpu marker).
 *
 * Semantics restated (file:line are relative to /root/reference):
 *   packed formats   LSB-first bit streams: value k of column n sits at bits [k*b, k*b+b) of the
 *                    stream qweight[:, n]; zero of (group, n) at bits [n*b, n*b+b) of qzeros[group, :]
 *                    (src/cu/unpack_weight_2_to_7.cu:53-66, :196-217, :256-281; src/dq_torch_ops.cc:31)
 *   dequant          sz  = RN16( RN16(z + add_zero_bias) * s )          (unpack_weight_2_to_7.cu:58-61, :284-286)
 *                    out = RN16( RN16(w) * s - sz )  [one rounding: hfma2] (unpack_weight_2_to_7.cu:72-75, :305, :313)
 *   gemv truth       y[m,n] = sum_k a[m,k] * out[k,n], accumulated in fp64 over the bit-exact
 *                    dequantised weights (the "fp32-accumulated reference" of the north star;
 *                    the shipped kernel src/cu/gemv_w4a16_pt.cu:35-145 uses fp16 chains of 4).
 * IEEE round-to-nearest-even is used for every fp16 rounding (the GPU's __hmul2/__hfma2);
 * the reference's CPU simulator rounds ties half-up (cpp_simulate.cc:47), hence "<= 1 ulp" there.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <math.h>

#define XO_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------- fp16 soft float */

static inline double f16_to_f64(uint16_t h) {
  const uint32_t sign = (h >> 15) & 1u, e = (h >> 10) & 0x1Fu, m = h & 0x3FFu;
  double v;
  if (e == 0)        v = ldexp((double)m, -24);                 /* zero / subnormal */
  else if (e == 31)  v = m ? NAN : INFINITY;
  else               v = ldexp((double)(m | 0x400u), (int)e - 25);
  return sign ? -v : v;
}

/* exact double -> fp16, round-to-nearest-even, single rounding */
static inline uint16_t f64_to_f16_rne(double d) {
  uint64_t bits; memcpy(&bits, &d, 8);
  const uint16_t sign = (uint16_t)((bits >> 48) & 0x8000u);
  const int e = (int)((bits >> 52) & 0x7FF);
  const uint64_t m = bits & 0xFFFFFFFFFFFFFull;
  if (e == 0x7FF) return (uint16_t)(sign | 0x7C00u | (m ? 0x200u : 0u));
  if (e == 0) return sign;                       /* double zero/subnormal -> +-0 */
  const int ue = e - 1023;                       /* unbiased exponent */
  uint64_t sig = m | (1ull << 52);               /* 53-bit significand */
  int shift;                                     /* bits to drop from sig */
  int he;                                        /* resulting biased half exponent (0 => subnormal) */
  if (ue >= -14) { shift = 42; he = ue + 15; }   /* normal half: keep 11 bits */
  else           { shift = 42 + (-14 - ue); he = 0; }
  if (shift > 63) return sign;                   /* far below half the smallest subnormal */
  uint64_t kept = sig >> shift;
  const uint64_t rem = sig & ((1ull << shift) - 1ull);
  const uint64_t half = 1ull << (shift - 1);
  if (rem > half || (rem == half && (kept & 1ull))) kept++;
  uint32_t out;
  if (he == 0) out = (uint32_t)kept;             /* subnormal; a carry into 0x400 is the smallest normal */
  else {
    out = ((uint32_t)he << 10) + (uint32_t)(kept - 0x400u); /* carry from mantissa bumps the exponent */
  }
  if (out >= 0x7C00u) out = 0x7C00u;             /* overflow -> inf */
  return (uint16_t)(sign | out);
}

XO_API uint16_t xo_f64_to_f16(double d) { return f64_to_f16_rne(d); }
XO_API double   xo_f16_to_f64(uint16_t h) { return f16_to_f64(h); }

/* ---------------------------------------------------------------- bit streams */

/* bits [pos, pos+b) of an LSB-first stream of 32-bit words with the given word stride.
 * Words at index >= nwords read as zero (ragged tail; the reference guards the same way,
 * unpack_weight_2_to_7.cu:236-239). */
static inline uint32_t stream_bits(const uint32_t* base, size_t stride, size_t nwords,
                                   uint64_t pos, int b) {
  const size_t w = (size_t)(pos >> 5);
  const int sh = (int)(pos & 31u);
  uint64_t lo = w < nwords ? base[w * stride] : 0u;
  uint64_t hi = (w + 1) < nwords ? base[(w + 1) * stride] : 0u;
  const uint64_t both = lo | (hi << 32);
  return (uint32_t)((both >> sh) & ((1u << b) - 1u));
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

/* w[k, n] for all k, n  (uint8 out[K*N]) */
XO_API void xo_unpack_qweight(const int32_t* qweight, int K, int N, int bits, uint8_t* out) {
  const size_t rows = (size_t)ceil_div(K * bits, 32);
  const uint32_t* q = (const uint32_t*)qweight;
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n)
      out[(size_t)k * N + n] = (uint8_t)stream_bits(q + n, (size_t)N, rows, (uint64_t)k * bits, bits);
}

/* raw z[g, n] (before add_zero_bias)  (uint8 out[G*N]) */
XO_API void xo_unpack_qzeros(const int32_t* qzeros, int G, int N, int bits, uint8_t* out) {
  const size_t zw = (size_t)ceil_div(N * bits, 32);
  const uint32_t* q = (const uint32_t*)qzeros;
  for (int g = 0; g < G; ++g)
    for (int n = 0; n < N; ++n)
      out[(size_t)g * N + n] = (uint8_t)stream_bits(q + (size_t)g * zw, 1, zw, (uint64_t)n * bits, bits);
}

/* ---------------------------------------------------------------- dequant */

/* out_f16[K, N] row-major.  scales_f16[G, N], qzeros[G, ceil(N*b/32)], G = ceil(K/groupsize).
 * Rows [k_begin, k_end) only: rows are independent, so callers may thread over row ranges
 * (the library itself is single-threaded and re-entrant). */
XO_API void xo_dequant_f16_rows(const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros,
                                uint16_t* out, int K, int N, int bits, int groupsize, int add_zero_bias,
                                int k_begin, int k_end) {
  const size_t rows = (size_t)ceil_div(K * bits, 32);
  const size_t zw = (size_t)ceil_div(N * bits, 32);
  const uint32_t* qw = (const uint32_t*)qweight;
  const uint32_t* qz = (const uint32_t*)qzeros;
  for (int k = k_begin; k < k_end && k < K; ++k) {
    const int g = k / groupsize;
    for (int n = 0; n < N; ++n) {
      const uint32_t w = stream_bits(qw + n, (size_t)N, rows, (uint64_t)k * bits, bits);
      const uint32_t z = stream_bits(qz + (size_t)g * zw, 1, zw, (uint64_t)n * bits, bits) + (uint32_t)add_zero_bias;
      const double s = f16_to_f64(scales[(size_t)g * N + n]);
      /* hmul2(half(z+bias), s): z+bias <= 256 is exact in fp16; product exact in double */
      const double sz = f16_to_f64(f64_to_f16_rne((double)z * s));
      /* hfma2(half(w), s, -sz): exact in double, one rounding */
      out[(size_t)k * N + n] = f64_to_f16_rne((double)w * s + (-sz));
    }
  }
}

XO_API void xo_dequant_f16(const int32_t* qweight, const uint16_t* scales, const int32_t* qzeros,
                           uint16_t* out, int K, int N, int bits, int groupsize, int add_zero_bias) {
  xo_dequant_f16_rows(qweight, scales, qzeros, out, K, N, bits, groupsize, add_zero_bias, 0, K);
}

/* ---------------------------------------------------------------- gemv truth */

/* y64[M, N] (fp64 accumulated over the bit-exact dequantised fp16 weights) and, if y16 != NULL,
 * its single RN16 rounding.  a_f16[M, K].  w_f16 is the [K, N] output of xo_dequant_f16. */
XO_API void xo_gemv_from_dq_cols(const uint16_t* a, const uint16_t* w_f16, double* y64, uint16_t* y16,
                                 int M, int K, int N, int n_begin, int n_end) {
  if (n_end > N) n_end = N;
  for (int n0 = n_begin; n0 < n_end; n0 += 64) {
    const int n1 = n0 + 64 < n_end ? n0 + 64 : n_end;
    for (int m = 0; m < M; ++m) {
      double acc[64];
      for (int j = 0; j < 64; ++j) acc[j] = 0.0;
      for (int k = 0; k < K; ++k) {
        const double av = f16_to_f64(a[(size_t)m * K + k]);
        const uint16_t* wr = w_f16 + (size_t)k * N;
        for (int n = n0; n < n1; ++n) acc[n - n0] += av * f16_to_f64(wr[n]);
      }
      for (int n = n0; n < n1; ++n) {
        y64[(size_t)m * N + n] = acc[n - n0];
        if (y16) y16[(size_t)m * N + n] = f64_to_f16_rne(acc[n - n0]);
      }
    }
  }
}

XO_API void xo_gemv_from_dq(const uint16_t* a, const uint16_t* w_f16, double* y64, uint16_t* y16,
                            int M, int K, int N) {
  xo_gemv_from_dq_cols(a, w_f16, y64, y16, M, K, N, 0, N);
}

/* fused convenience: dequant (scratch w_f16[K*N] supplied by the caller) then truth gemv */
XO_API void xo_gemv_f16(const uint16_t* a, const int32_t* qweight, const uint16_t* scales,
                        const int32_t* qzeros, double* y64, uint16_t* y16, uint16_t* w_scratch,
                        int M, int K, int N, int bits, int groupsize, int add_zero_bias) {
  xo_dequant_f16(qweight, scales, qzeros, w_scratch, K, N, bits, groupsize, add_zero_bias);
  xo_gemv_from_dq(a, w_scratch, y64, y16, M, K, N);
}

/*
 * The reference's SHIPPED gemv arithmetic (src/cu/gemv_w4a16_pt.cu:67-143), restated for
 * information (how far the reference itself sits from the truth): w' = hfma2(s, w, -hmul(z, s));
 * 4-deep fp16 hfma2 chain over (k, k+1) pairs with the activations; fp32 partial sums per
 * K-slab thread; fp32 cross-slab sum; RN16.  add_zero_bias is applied to every group (the
 * intended semantics; the shipped kernel drops it after the first group of a slab).
 * bits = 4 only.  Summation order across slabs: ascending slab index (the warp shuffle tree of
 * the kernel is not reproduced; this is an fp32 reassociation, not part of the parity bar).
 */
XO_API void xo_gemv_w4_ref_arith(const uint16_t* a, const int32_t* qweight, const uint16_t* scales,
                                 const int32_t* qzeros, uint16_t* y16,
                                 int M, int K, int N, int groupsize, int add_zero_bias) {
  const size_t zw = (size_t)ceil_div(N * 4, 32);
  const uint32_t* qw = (const uint32_t*)qweight;
  const uint32_t* qz = (const uint32_t*)qzeros;
  const int block_k = ((K + 31) / 32 + 7) / 8 * 8;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float total = 0.f;
      for (int y0 = 0; y0 < K; y0 += block_k) {
        float sum = 0.f;
        for (int kc = y0; kc < y0 + block_k && kc < K; kc += 8) {
          const int g = kc / groupsize;
          const double s = f16_to_f64(scales[(size_t)g * N + n]);
          const uint32_t z = ((qz[(size_t)g * zw + (size_t)(n >> 3)] >> (4 * (n & 7))) & 0xFu) + (uint32_t)add_zero_bias;
          const double zs = f16_to_f64(f64_to_f16_rne((double)z * s));
          const uint32_t word = qw[(size_t)(kc >> 3) * N + n];
          double rx = 0.0, ry = 0.0;  /* the two halves of res2, kept as exact fp16 values */
          for (int j = 0; j < 4; ++j) {
            const double w0 = f16_to_f64(f64_to_f16_rne((double)((word >> (8 * j)) & 0xFu) * s + (-zs)));
            const double w1 = f16_to_f64(f64_to_f16_rne((double)((word >> (8 * j + 4)) & 0xFu) * s + (-zs)));
            const int k0 = kc + 2 * j;
            const double a0 = k0 < K ? f16_to_f64(a[(size_t)m * K + k0]) : 0.0;
            const double a1 = (k0 + 1) < K ? f16_to_f64(a[(size_t)m * K + k0 + 1]) : 0.0;
            rx = f16_to_f64(f64_to_f16_rne(a0 * w0 + rx));
            ry = f16_to_f64(f64_to_f16_rne(a1 * w1 + ry));
          }
          sum += (float)rx + (float)ry;
        }
        total += sum;
      }
      y16[(size_t)m * N + n] = f64_to_f16_rne((double)total);
    }
}

/* ---------------------------------------------------------------- packer (inverse of the unpackers) */

/* w_u8[K, N] -> qweight[ceil(K*b/32), N]; the destination must be zero-initialised. */
XO_API void xo_pack_qweight(const uint8_t* w, int K, int N, int bits, int32_t* qweight) {
  uint32_t* q = (uint32_t*)qweight;
  const uint32_t mask = (1u << bits) - 1u;
  for (int k = 0; k < K; ++k)
    for (int n = 0; n < N; ++n) {
      const uint64_t pos = (uint64_t)k * bits;
      const size_t wi = (size_t)(pos >> 5);
      const int sh = (int)(pos & 31u);
      const uint64_t v = (uint64_t)(w[(size_t)k * N + n] & mask) << sh;
      q[wi * N + n] |= (uint32_t)v;
      if (sh + bits > 32) q[(wi + 1) * N + n] |= (uint32_t)(v >> 32);
    }
}

/* z_u8[G, N] (raw, i.e. already minus add_zero_bias) -> qzeros[G, ceil(N*b/32)]; zero-initialised dst. */
XO_API void xo_pack_qzeros(const uint8_t* z, int G, int N, int bits, int32_t* qzeros) {
  uint32_t* q = (uint32_t*)qzeros;
  const size_t zw = (size_t)ceil_div(N * bits, 32);
  const uint32_t mask = (1u << bits) - 1u;
  for (int g = 0; g < G; ++g)
    for (int n = 0; n < N; ++n) {
      const uint64_t pos = (uint64_t)n * bits;
      const size_t wi = (size_t)(pos >> 5);
      const int sh = (int)(pos & 31u);
      const uint64_t v = (uint64_t)(z[(size_t)g * N + n] & mask) << sh;
      q[(size_t)g * zw + wi] |= (uint32_t)v;
      if (sh + bits > 32) q[(size_t)g * zw + wi + 1] |= (uint32_t)(v >> 32);
    }
}
