// butterflies.cuh -- register-level DFT butterflies (natural-order in, natural-order out).
//
// INV = true computes X[m] = sum_n x[n] e^{+2 pi i n m / R} (the inverse-transform sign used by
// the k-space -> image path); INV = false the forward sign.  No scaling is applied here.
#pragma once
#include <utility>
#include "common.cuh"
#include "dft_consts.h"

namespace mriacl {

template <class F, int... I>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}
// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>)
template <int N, class F> __device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(f, std::make_integer_sequence<int, N>{});
}

#define MRIACL_SQRT1_2 0.70710678118654752440f

template <bool INV> __device__ __forceinline__ void radix2(cf& a, cf& b) {
  cf t = a; a = cadd(t, b); b = csub(t, b);
}

// 4-point DFT of (a0,a1,a2,a3) in place
template <bool INV> __device__ __forceinline__ void radix4(cf& a0, cf& a1, cf& a2, cf& a3) {
  cf t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_i<INV>(csub(a1, a3));
  a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// 8-point DFT, v[0..7] in place, natural order
template <bool INV> __device__ __forceinline__ void radix8(cf* v) {
  // even / odd 4-point transforms
  radix4<INV>(v[0], v[2], v[4], v[6]);   // E[0..3] in v[0],v[2],v[4],v[6]
  radix4<INV>(v[1], v[3], v[5], v[7]);   // O[0..3] in v[1],v[3],v[5],v[7]
  // X[m], X[m+4] = E[m] +- O[m] w8^m with w8 = (1 +- i)/sqrt 2:  O1 w8 = c (O1 +- i O1),  O3 w8^3 = c (+-i O3 - O3);
  // the scale c rides on the final packed FMA
  const float c = MRIACL_SQRT1_2;
  const cf o0 = v[1];
  const cf t1 = cadd(v[3], mul_i<INV>(v[3]));
  const cf o2 = mul_i<INV>(v[5]);
  const cf t3 = csub(mul_i<INV>(v[7]), v[7]);
  const cf e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = pk_fma(t1, bc(c), e1); v[5] = pk_fma(t1, bc(-c), e1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = pk_fma(t3, bc(c), e3); v[7] = pk_fma(t3, bc(-c), e3);
}

// 5-point DFT in place
template <bool INV> __device__ __forceinline__ void radix5(cf& a0, cf& a1, cf& a2, cf& a3, cf& a4) {
  constexpr float c1 = DftConsts<5>::c(1), c2 = DftConsts<5>::c(2);
  constexpr float s1 = DftConsts<5>::s(1), s2 = DftConsts<5>::s(2);
  const cf p1 = cadd(a1, a4), p2 = cadd(a2, a3), d1 = csub(a1, a4), d2 = csub(a2, a3);
  const cf r1 = pk_fma(p2, bc(c2), pk_fma(p1, bc(c1), a0));
  const cf r2 = pk_fma(p2, bc(c1), pk_fma(p1, bc(c2), a0));
  const cf i1 = pk_fma(d2, bc(s2), pk_mul(d1, bc(s1)));
  const cf i2 = pk_fma(d2, bc(-s1), pk_mul(d1, bc(s2)));
  a0 = cadd(a0, cadd(p1, p2));
  const cf j1 = mul_i<INV>(i1), j2 = mul_i<INV>(i2);
  a1 = cadd(r1, j1); a4 = csub(r1, j1);
  a2 = cadd(r2, j2); a3 = csub(r2, j2);
}

// 10-point DFT, v[0..9] in place, natural order (2 x radix-5 + radix-2 recombination)
template <bool INV> __device__ __forceinline__ void radix10(cf* v) {
  radix5<INV>(v[0], v[2], v[4], v[6], v[8]);   // E[m] in v[2m]
  radix5<INV>(v[1], v[3], v[5], v[7], v[9]);   // O[m] in v[2m+1]
  cf e[5] = {v[0], v[2], v[4], v[6], v[8]};
  cf o[5];
  o[0] = v[1];
  static_for<4>([&](auto mm) {
    constexpr int m = mm.value + 1;
    // w10^m = exp(+-2 pi i m / 10) = (cos, sin)(2 pi 2m/20); take it from the radix-5 table:
    // cos(2 pi m/10) = -cos(2 pi (m+5)/10) and 2 pi m/10 = 2 pi (m * 3 mod 5)... keep literals instead
    constexpr float cw[5] = {1.0f, 0.80901699437494742410f, 0.30901699437494742410f,
                             -0.30901699437494742410f, -0.80901699437494742410f};
    constexpr float sw[5] = {0.0f, 0.58778525229247312917f, 0.95105651629515357212f,
                             0.95105651629515357212f, 0.58778525229247312917f};
    o[m] = cmul_k(v[2 * m + 1], cw[m], INV ? sw[m] : -sw[m]);
  });
  static_for<5>([&](auto mm) {
    constexpr int m = mm.value;
    v[m] = cadd(e[m], o[m]);
    v[m + 5] = csub(e[m], o[m]);
  });
}

// 16-point DFT, v[0..15] in place, natural order (4 x 4)
template <bool INV> __device__ __forceinline__ void fft16(cf* v) {
  // step 1: for each b, 4-point DFT over a of v[4a + b]  -> u[b][c] stored at v[4c + b]
  static_for<4>([&](auto bb) { constexpr int b = bb.value; radix4<INV>(v[b], v[4 + b], v[8 + b], v[12 + b]); });
  // step 2: u[b][c] *= w16^{b c}
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin(pi/8)
  constexpr float h = MRIACL_SQRT1_2;
  constexpr float cw[10] = {1.f, c1, h, s1, 0.f, -s1, -h, -c1, -1.f, -c1};
  constexpr float sw[10] = {0.f, s1, h, c1, 1.f, c1, h, s1, 0.f, -s1};
  static_for<3>([&](auto bb) {
    constexpr int b = bb.value + 1;
    static_for<3>([&](auto cc) {
      constexpr int c = cc.value + 1;
      constexpr int e = b * c;  // 1..9
      constexpr float wc = cw[e], ws = INV ? sw[e] : -sw[e];
      if constexpr (wc == 0.f) v[4 * c + b] = cscale(mul_i<true>(v[4 * c + b]), ws);                 // +-i
      else if constexpr (wc == ws) v[4 * c + b] = cscale(cadd(v[4 * c + b], mul_i<true>(v[4 * c + b])), wc);
      else v[4 * c + b] = cmul_k(v[4 * c + b], wc, ws);
    });
  });
  // step 3: for each c, 4-point DFT over b of v[4c + b] -> X[c + 4d] at v[4c + d]
  static_for<4>([&](auto cc) { constexpr int c = cc.value; radix4<INV>(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]); });
  // v[4c + d] = X[c + 4d]  -> transpose the 4x4 to natural order
  static_for<4>([&](auto cc) {
    constexpr int c = cc.value;
    static_for<4>([&](auto dd) {
      constexpr int d = dd.value;
      if constexpr (d > c) { cf t = v[4 * c + d]; v[4 * c + d] = v[4 * d + c]; v[4 * d + c] = t; }
    });
  });
}

// 3-point DFT in place
template <bool INV> __device__ __forceinline__ void radix3(cf& a0, cf& a1, cf& a2) {
  constexpr float hs = 0.86602540378443864676f;   // sin(2 pi / 3)
  const cf t = cadd(a1, a2);
  const cf m = pk_fma(t, bc(-0.5f), a0);
  const cf d = cscale(csub(a1, a2), hs);
  const cf j = mul_i<INV>(d);
  a0 = cadd(a0, t);
  a1 = cadd(m, j);
  a2 = csub(m, j);
}

// 12-point DFT, v[0..11] in place, natural order (n = 3a + b, k = c + 4d: four-point transforms over a, twiddles
// w12^{bc}, three-point transforms over b)
template <bool INV> __device__ __forceinline__ void fft12(cf* v) {
  static_for<3>([&](auto bb) { constexpr int b = bb.value; radix4<INV>(v[b], v[3 + b], v[6 + b], v[9 + b]); });   // u[b][c] at v[3c + b]
  constexpr float h3 = 0.86602540378443864676f;
  constexpr float cw[7] = {1.f, h3, 0.5f, 0.f, -0.5f, -h3, -1.f};
  constexpr float sw[7] = {0.f, 0.5f, h3, 1.f, h3, 0.5f, 0.f};
  static_for<2>([&](auto bb) {
    constexpr int b = bb.value + 1;
    static_for<3>([&](auto cc) {
      constexpr int c = cc.value + 1;
      constexpr int e = b * c;      // 1, 2, 3, 2, 4, 6
      constexpr float wc = cw[e], ws = INV ? sw[e] : -sw[e];
      if constexpr (e == 6) v[3 * c + b] = cneg(v[3 * c + b]);
      else if constexpr (e == 3) v[3 * c + b] = cscale(mul_i<true>(v[3 * c + b]), ws);
      else v[3 * c + b] = cmul_k(v[3 * c + b], wc, ws);
    });
  });
  static_for<4>([&](auto cc) { constexpr int c = cc.value; radix3<INV>(v[3 * c], v[3 * c + 1], v[3 * c + 2]); });   // X[c + 4d] at v[3c + d]
  cf t[12];
  static_for<12>([&](auto ii) { t[ii.value] = v[ii.value]; });
  static_for<4>([&](auto cc) {
    constexpr int c = cc.value;
    static_for<3>([&](auto dd) { constexpr int d = dd.value; v[c + 4 * d] = t[3 * c + d]; });
  });
}

// Q-point register FFT of the row pass's second stage
template <int Q, bool INV> __device__ __forceinline__ void fft_q(cf* v) {
  static_assert(Q == 16 || Q == 12, "second-stage lengths: 16 (368 = 23 x 16) and 12 (372 = 31 x 12)");
  if constexpr (Q == 16) fft16<INV>(v); else fft12<INV>(v);
}

// P-point DFT for odd P by the symmetric direct method:
//   a_n = x[n] + x[P-n], b_n = x[n] - x[P-n]  (n = 1..(P-1)/2)
//   X[k], X[P-k] = x0 + sum a_n cos(2 pi n k/P)  +-  i sum b_n sin(2 pi n k/P)
// ~ (P-1)^2 real FMAs with immediate constants.  Results are handed to emit(k, X[k]) as they
// are produced so the caller can twiddle and store without keeping all P outputs live.
// The _part form computes only the output pairs k in [KLO, KHI) (and X[0] when WITH0), so one
// transform can be shared between two warps.
template <int P, bool INV, int KLO, int KHI, bool WITH0, class X, class Emit>
__device__ __forceinline__ void dft_odd_sym_part(const X& x, Emit&& emit) {
  constexpr int Hh = (P - 1) / 2;
  constexpr int NK = KHI - KLO;
  static_assert(KLO >= 1 && KHI <= Hh + 1 && KLO <= KHI, "pair range");
  // streaming order: each symmetric couple (x[n], x[P-n]) is combined once and fed to all NK output pairs, so only the
  // 2 NK running sums (and X[0]'s) stay live -- not the Hh sums and differences
  const cf x0 = x[0];
  cf A[NK > 0 ? NK : 1], B[NK > 0 ? NK : 1];
  cf s0 = x0;
  static_for<NK>([&](auto kk) { A[kk.value] = x0; });
  static_for<Hh>([&](auto nn) {
    constexpr int n = nn.value + 1;
    const cf a = cadd(x[n], x[P - n]), b = csub(x[n], x[P - n]);
    if constexpr (WITH0) s0 = cadd(s0, a);
    static_for<NK>([&](auto kk) {
      constexpr int k = kk.value + KLO;
      constexpr float c = DftConsts<P>::c((n * k) % P);
      constexpr float s = DftConsts<P>::s((n * k) % P);
      A[kk.value] = pk_fma(a, bc(c), A[kk.value]);
      if constexpr (n == 1) B[kk.value] = pk_mul(b, bc(s)); else B[kk.value] = pk_fma(b, bc(s), B[kk.value]);
    });
  });
  if constexpr (WITH0) emit(std::integral_constant<int, 0>{}, s0);
  static_for<NK>([&](auto kk) {
    constexpr int k = kk.value + KLO;
    // A + iB and A - iB
    const cf iB = mul_i<true>(B[kk.value]);
    const cf plus = cadd(A[kk.value], iB), minus = csub(A[kk.value], iB);
    emit(std::integral_constant<int, k>{}, INV ? plus : minus);
    emit(std::integral_constant<int, P - k>{}, INV ? minus : plus);
  });
}

template <int P, bool INV, class X, class Emit>
__device__ __forceinline__ void dft_odd_sym(const X& x, Emit&& emit) {
  dft_odd_sym_part<P, INV, 1, (P - 1) / 2 + 1, true>(x, emit);
}

}  // namespace mriacl
