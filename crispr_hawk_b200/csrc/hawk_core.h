// hawk_core.h -- bit-level building blocks shared by the CUDA kernels and the
// host-side self-check harness (hostcheck.cpp). Everything here is
// __host__ __device__ so the exact code the kernels run can be exercised on a
// CPU against the oracle before it ever reaches a GPU.
//
// Layout recap (include/hawkscan.h): one chunk = 32 consecutive base slots;
// planes q[chunk] = {A,C,G,T} (bit i of each word belongs to slot 32*chunk+i),
// v[chunk] = lower-case ("variant base") bits.
#pragma once
#include <stdint.h>

#include "../../include/hawkscan.h"

#if defined(__CUDACC__)
#define HAWK_HD __host__ __device__ __forceinline__
#define HAWK_UNROLL _Pragma("unroll")
#else
#define HAWK_HD inline
#define HAWK_UNROLL
#endif

namespace hawk {

struct alignas(16) Planes {
  uint32_t a, c, g, t;
};

// ---- ASCII -> 4-bit IUPAC mask (encoder.py:18-34) -------------------------------
// entry: bits 0-3 nibble, bit 4 lower-case, 0x80 invalid (non-IUPAC, non-NUL)
HAWK_HD uint8_t iupac_entry(uint8_t ch) {
  if (ch == 0) return 0;  // unused slot
  uint8_t lower = (ch >= 'a' && ch <= 'z') ? 0x10 : 0;
  uint8_t up = lower ? (uint8_t)(ch - 32) : ch;
  uint8_t n;
  switch (up) {
    case 'A': n = 1; break;
    case 'C': n = 2; break;
    case 'G': n = 4; break;
    case 'T': n = 8; break;
    case 'R': n = 5; break;
    case 'Y': n = 10; break;
    case 'S': n = 6; break;
    case 'W': n = 9; break;
    case 'K': n = 12; break;
    case 'M': n = 3; break;
    case 'B': n = 14; break;
    case 'D': n = 13; break;
    case 'H': n = 11; break;
    case 'V': n = 7; break;
    case 'N': n = 15; break;
    default: return 0x80;
  }
  return (uint8_t)(n | lower);
}

// ---- K1: 32 ASCII bytes -> bit-sliced planes, no table lookups ---------------------
// (1) a bit-matrix transpose turns the characters into bit planes b0..b7 (bit i of plane k =
//     bit k of character i);
// (2) the letter's low five bits index 32-entry truth tables, one per output plane,
//     evaluated for 32 characters at once with 3-input boolean ops (LOP3 on the GPU).
// Two forms with the same results: pack_chunk, the plain one (an 8x8 transpose per 8 bytes by
// delta swaps, exact tables by Shannon expansion over b4, b3) -- round 1's kernel code, kept as
// the form the tests compare with --, and pack_chunk_v3 further down, which the kernels run.

// 3-input boolean function selected by the 8-bit truth table IMM (index a<<2 | b<<1 | c)
template <int IMM>
HAWK_HD uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(d) : "r"(a), "r"(b), "r"(c), "n"(IMM));
  return d;
#else
  uint32_t r = 0;
  for (int i = 0; i < 8; ++i)
    if (IMM & (1 << i)) r |= ((i & 4) ? a : ~a) & ((i & 2) ? b : ~b) & ((i & 1) ? c : ~c);
  return r;
#endif
}

HAWK_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(a, b, sel);
#else
  uint64_t both = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) r |= (uint32_t)((both >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
  return r;
#endif
}

// 8x8 bit-matrix transpose of the eight bytes held in (lo, hi): afterwards byte k holds bit k
// of the eight input bytes (bit i of it comes from input byte i). Three delta swaps; the first
// two stay inside each 32-bit half. t and its shifted copy never overlap, so "t ^ (t << s)"
// is the multiply t * (1 + 2^s): that moves a third of the work from the ALU pipe (shifts,
// LOP3) to the otherwise idle FMA pipe (IMAD) -- K1 is bound by the ALU pipe, not by HBM.
// x >> S through the multiplier (high half of x * 2^(32-S)) on the GPU
template <int S>
HAWK_HD uint32_t shr_mul(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __umulhi(x, 1u << (32 - S));
#else
  return x >> S;
#endif
}

HAWK_HD void transpose8x8(uint32_t& lo, uint32_t& hi) {
  uint32_t t;
  t = (lo ^ shr_mul<7>(lo)) & 0x00AA00AAu;
  lo ^= t * 129u;
  t = (hi ^ shr_mul<7>(hi)) & 0x00AA00AAu;
  hi ^= t * 129u;
  t = (lo ^ shr_mul<14>(lo)) & 0x0000CCCCu;
  lo ^= t * 16385u;
  t = (hi ^ shr_mul<14>(hi)) & 0x0000CCCCu;
  hi ^= t * 16385u;
  t = (lo ^ (hi * 16u)) & 0xF0F0F0F0u;
  lo ^= t;
  hi ^= shr_mul<4>(t);
}

// truth table over the letter number (ch & 31): bit L set <=> the IUPAC code of letter L
// contains base `bit` (encoder.py:18-34)
constexpr uint32_t letter_table(int bit) {
  uint32_t m = 0;
  const char letters[] = "ACGTRYSWKMBDHVN";
  const int codes[] = {1, 2, 4, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};
  for (int i = 0; i < 15; ++i)
    if (codes[i] & (1 << bit)) m |= 1u << (letters[i] & 31);
  return m;
}

// f(b4..b0) for 32 characters at once, f given as a 32-entry truth table
template <uint32_t TABLE>
HAWK_HD uint32_t table5(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3, uint32_t b4) {
  uint32_t g0 = lop3<(int)(TABLE & 0xFF)>(b2, b1, b0);
  uint32_t g1 = lop3<(int)((TABLE >> 8) & 0xFF)>(b2, b1, b0);
  uint32_t g2 = lop3<(int)((TABLE >> 16) & 0xFF)>(b2, b1, b0);
  uint32_t g3 = lop3<(int)((TABLE >> 24) & 0xFF)>(b2, b1, b0);
  uint32_t lo = lop3<0xCA>(b3, g1, g0);  // b3 ? g1 : g0
  uint32_t hi = lop3<0xCA>(b3, g3, g2);
  return lop3<0xCA>(b4, hi, lo);
}

struct PackedChunk {
  uint32_t a, c, g, t, v, invalid;
};

// 32 ASCII bytes (8 little-endian words) -> plane words. NUL bytes (unused slots) give
// all-zero bits; any other non-IUPAC byte sets its `invalid` bit.
HAWK_HD PackedChunk pack_chunk(const uint32_t* words) {
  uint32_t lo[4], hi[4];
  HAWK_UNROLL
  for (int g = 0; g < 4; ++g) {
    lo[g] = words[2 * g];      // -> bytes: bit planes 0..3 of characters 8g .. 8g+7
    hi[g] = words[2 * g + 1];  // -> bit planes 4..7
    transpose8x8(lo[g], hi[g]);
  }
  uint32_t b[8];
  HAWK_UNROLL
  for (int k = 0; k < 4; ++k) {
    const uint32_t sel = (uint32_t)k | ((uint32_t)(4 + k) << 4);  // byte k of a, byte k of b
    b[k] = byte_perm(byte_perm(lo[0], lo[1], sel), byte_perm(lo[2], lo[3], sel), 0x5410);
    b[4 + k] = byte_perm(byte_perm(hi[0], hi[1], sel), byte_perm(hi[2], hi[3], sel), 0x5410);
  }
  const uint32_t alpha = ~b[7] & b[6];  // 0x40..0x7F; b5 = lower-case
  PackedChunk o;
  o.a = alpha & table5<letter_table(0)>(b[0], b[1], b[2], b[3], b[4]);
  o.c = alpha & table5<letter_table(1)>(b[0], b[1], b[2], b[3], b[4]);
  o.g = alpha & table5<letter_table(2)>(b[0], b[1], b[2], b[3], b[4]);
  o.t = alpha & table5<letter_table(3)>(b[0], b[1], b[2], b[3], b[4]);
  const uint32_t valid = o.a | o.c | o.g | o.t;
  const uint32_t nul = ~(b[0] | b[1] | b[2] | b[3] | b[4] | b[5] | b[6] | b[7]);
  o.v = valid & b[5];
  o.invalid = ~(valid | nul);
  return o;
}

// 4x4 byte-matrix transpose of four words (row = word, column = byte): eight byte permutes
HAWK_HD void transpose4x4_bytes(uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  const uint32_t t0 = byte_perm(w0, w1, 0x5140), t1 = byte_perm(w2, w3, 0x5140);
  const uint32_t t2 = byte_perm(w0, w1, 0x7362), t3 = byte_perm(w2, w3, 0x7362);
  w0 = byte_perm(t0, t1, 0x5410);
  w1 = byte_perm(t0, t1, 0x7632);
  w2 = byte_perm(t2, t3, 0x5410);
  w3 = byte_perm(t2, t3, 0x7632);
}

// ---- pack_chunk_v3: the same planes with two thirds of the instructions (what the kernels run;
// pack_chunk above stays as the plain form the tests compare it with) --------------------------
// (1) Transpose. Two 4x4 byte transposes first put characters {m, m + 8, m + 16, m + 24} into word m,
//     so a bit's address is (word m = i mod 8, position 8 (i div 8) + j) for bit j of character i.
//     The wanted address is (word j, position i): exchange word-index bit k with position bit k for
//     k = 0, 1, 2 -- three stages of four word pairs, each pair two shifts (through the multiplier)
//     and two bit selects, instead of delta swaps inside 8-byte groups followed by byte gathers:
//     16 permutes + 24 shifts + 24 selects against 16 + 40 + 44.
// (2) Truth tables. Only 16 of the 32 letter numbers matter for the planes -- the 15 IUPAC letters
//     and 0 (NUL, zero planes); any other byte makes the whole batch invalid -- so each plane is the
//     smallest 3-input-gate circuit that agrees on those 16 (tools/lop3_search.py: 3 + 2 + 3 + 3
//     gates instead of 4 x 7); validity is its own exact 5-input function.
// Same result as pack_chunk on every chunk of NUL / IUPAC bytes and the same `invalid` word on
// every input (tests/test_hostcheck.py).
template <int S>
HAWK_HD uint32_t shl_mul(uint32_t x) {
  return x * (1u << S);
}

// bits of `a` at positions whose bit S is set <-> bits of `b` S positions below (M: positions with
// bit S clear)
template <int S, uint32_t M>
HAWK_HD void bit_exchange(uint32_t& a, uint32_t& b) {
  const uint32_t up = shl_mul<S>(b), down = shr_mul<S>(a);
  a = lop3<0xCA>(M, a, up);    // M ? a : b << S
  b = lop3<0xCA>(M, down, b);  // M ? a >> S : b
}

constexpr uint32_t valid_letter_table() { return letter_table(0) | letter_table(1) | letter_table(2) | letter_table(3); }

HAWK_HD PackedChunk pack_chunk_v3(const uint32_t* words) {
  uint32_t b0 = words[0], b1 = words[2], b2 = words[4], b3 = words[6];
  uint32_t b4 = words[1], b5 = words[3], b6 = words[5], b7 = words[7];
  transpose4x4_bytes(b0, b1, b2, b3);  // word m: characters m, m + 8, m + 16, m + 24
  transpose4x4_bytes(b4, b5, b6, b7);
  bit_exchange<1, 0x55555555u>(b0, b1);
  bit_exchange<1, 0x55555555u>(b2, b3);
  bit_exchange<1, 0x55555555u>(b4, b5);
  bit_exchange<1, 0x55555555u>(b6, b7);
  bit_exchange<2, 0x33333333u>(b0, b2);
  bit_exchange<2, 0x33333333u>(b1, b3);
  bit_exchange<2, 0x33333333u>(b4, b6);
  bit_exchange<2, 0x33333333u>(b5, b7);
  bit_exchange<4, 0x0F0F0F0Fu>(b0, b4);
  bit_exchange<4, 0x0F0F0F0Fu>(b1, b5);
  bit_exchange<4, 0x0F0F0F0Fu>(b2, b6);
  bit_exchange<4, 0x0F0F0F0Fu>(b3, b7);  // b_j = bit plane j of the 32 characters
  PackedChunk o;
  o.a = lop3<0xC2>(lop3<29>(b0, b1, b2), lop3<22>(b0, b3, b4), b2);
  o.c = lop3<0x6A>(lop3<57>(b0, b2, b4), b1, b3);
  o.g = lop3<0x16>(lop3<3>(b0, b1, b2), lop3<107>(b0, b2, b4), b3);
  o.t = lop3<0x69>(lop3<35>(b0, b1, b2), lop3<126>(b0, b2, b4), b3);
  o.v = b5;
  const uint32_t letter = table5<valid_letter_table()>(b0, b1, b2, b3, b4);
  const uint32_t any7 = lop3<0xFE>(lop3<0xFE>(b0, b1, b2), lop3<0xFE>(b3, b4, b5), b6);
  const uint32_t ok = lop3<0x40>(letter, b6, b7);  // an IUPAC letter: 0x40..0x7F with a valid number
  o.invalid = lop3<0x54>(any7, b7, ok);            // (any7 | b7) & ~ok: not NUL, not a letter
  return o;
}

// ---- planes -> text: 32 window characters from plane bits (extract_guide_sequence, :134-160) ----
// The way back of pack_chunk_v3. (1) The ASCII bit planes of the letters "?ACMGRSVTWYHKDBN"[nibble]
// as boolean functions of the four nibble planes (Shannon expansion over T: three 3-input gates per
// bit), bit 5 also set where the case plane is (lower-case = variant base). (2) The exchange
// stages of the transpose are involutions and commute, and so are the byte transposes: running
// them again turns the eight bit planes into the 32 characters. ~90 instructions for 32
// characters; the per-character table walk it replaces (spread four plane bits to bytes, build a
// permute selector, two table permutes, merge) took ~35 per 4.
constexpr uint32_t letter_bit_table(int bit, int t) {
  uint32_t m = 0;
  const char letters[] = "?ACMGRSVTWYHKDBN";
  for (int a = 0; a < 2; ++a)
    for (int c = 0; c < 2; ++c)
      for (int g = 0; g < 2; ++g)
        if ((letters[a + 2 * c + 4 * g + 8 * t] >> bit) & 1) m |= 1u << (a * 4 + c * 2 + g);
  return m;
}

template <int BIT>
HAWK_HD uint32_t ascii_plane(uint32_t a, uint32_t c, uint32_t g, uint32_t t) {
  return lop3<0xCA>(t, lop3<(int)letter_bit_table(BIT, 1)>(a, c, g), lop3<(int)letter_bit_table(BIT, 0)>(a, c, g));
}

// w[0..7]: characters 0..31 (little-endian words); `valid`: bit i set = character i exists, the
// others come out as zero bytes
HAWK_HD void planes_to_chars32(uint32_t pa, uint32_t pc, uint32_t pg, uint32_t pt, uint32_t pv, uint32_t valid,
                               uint32_t* w) {
  uint32_t b0 = ascii_plane<0>(pa, pc, pg, pt), b1 = ascii_plane<1>(pa, pc, pg, pt), b2 = ascii_plane<2>(pa, pc, pg, pt);
  uint32_t b3 = ascii_plane<3>(pa, pc, pg, pt), b4 = ascii_plane<4>(pa, pc, pg, pt);
  uint32_t b5 = ascii_plane<5>(pa, pc, pg, pt) | pv, b6 = ascii_plane<6>(pa, pc, pg, pt), b7 = 0u;
  if (valid != 0xFFFFFFFFu) {
    b0 &= valid, b1 &= valid, b2 &= valid, b3 &= valid, b4 &= valid, b5 &= valid, b6 &= valid;
  }
  bit_exchange<4, 0x0F0F0F0Fu>(b0, b4);
  bit_exchange<4, 0x0F0F0F0Fu>(b1, b5);
  bit_exchange<4, 0x0F0F0F0Fu>(b2, b6);
  bit_exchange<4, 0x0F0F0F0Fu>(b3, b7);
  bit_exchange<2, 0x33333333u>(b0, b2);
  bit_exchange<2, 0x33333333u>(b1, b3);
  bit_exchange<2, 0x33333333u>(b4, b6);
  bit_exchange<2, 0x33333333u>(b5, b7);
  bit_exchange<1, 0x55555555u>(b0, b1);
  bit_exchange<1, 0x55555555u>(b2, b3);
  bit_exchange<1, 0x55555555u>(b4, b5);
  bit_exchange<1, 0x55555555u>(b6, b7);
  transpose4x4_bytes(b0, b1, b2, b3);
  transpose4x4_bytes(b4, b5, b6, b7);
  w[0] = b0, w[2] = b1, w[4] = b2, w[6] = b3;
  w[1] = b4, w[3] = b5, w[5] = b6, w[7] = b7;
}

// nibble -> IUPAC letter (inverse of the table above), upper-case
HAWK_HD char nibble_letter(uint32_t n) {
  // "?ACMGRSVTWYHKDBN" as two 64-bit immediates (no local array, no stack frame)
  n &= 15;
  const uint64_t t0 = ((uint64_t)'?') | ((uint64_t)'A' << 8) | ((uint64_t)'C' << 16) |
                      ((uint64_t)'M' << 24) | ((uint64_t)'G' << 32) | ((uint64_t)'R' << 40) |
                      ((uint64_t)'S' << 48) | ((uint64_t)'V' << 56);
  const uint64_t t1 = ((uint64_t)'T') | ((uint64_t)'W' << 8) | ((uint64_t)'Y' << 16) |
                      ((uint64_t)'H' << 24) | ((uint64_t)'K' << 32) | ((uint64_t)'D' << 40) |
                      ((uint64_t)'B' << 48) | ((uint64_t)'N' << 56);
  return (char)(((n & 8) ? t1 : t0) >> ((n & 7) * 8));
}

HAWK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, sh);
#else
  sh &= 31;
  return sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
#endif
}

HAWK_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// bits i of a 32-slot chunk starting at p0 with lo <= p0 + i < hi
HAWK_HD uint32_t interval_mask(int32_t lo, int32_t hi, int32_t p0) {
  int32_t l = lo - p0, h = hi - p0;
  if (l < 0) l = 0;
  if (h > 32) h = 32;
  if (h <= l) return 0u;
  uint32_t upper = (h == 32) ? 0xFFFFFFFFu : ((1u << h) - 1u);
  uint32_t lower = (1u << l) - 1u;  // l < 32 here
  return upper & ~lower;
}

// plane union selected by an IUPAC nibble: bit i set <=> nibble(slot i) & code != 0
HAWK_HD uint32_t select_planes(const Planes& p, uint32_t code) {
  uint32_t r = 0;
  if (code & 1) r |= p.a;
  if (code & 2) r |= p.c;
  if (code & 4) r |= p.g;
  if (code & 8) r |= p.t;
  return r;
}

// search_guides.py:32-46 (match) for the 32 positions of one chunk at once, both
// patterns together: bit i of m[s] <=> for every k < P: pat[s][k] & nibble(p0 + i + k) != 0.
// `cur` / `nxt` are the planes of the chunk and of the following chunk (P <= 16 never
// reaches further). sel[s][k] = one all-ones / all-zeros word per base of pattern nibble k
// (branch-free AND-mask plane select); the shifted planes are shared by the two patterns.
// Pattern nibbles equal to 15 (N) match every real base and are skipped (skip bit k of
// skip[s]).
struct PamSelect {
  uint32_t sel[2][HAWK_MAX_PAM][4];
  uint32_t skip[2];
};

template <int P>
HAWK_HD void match_fixed(const Planes& cur, const Planes& nxt, const PamSelect& S, uint32_t m[2]) {
  m[0] = m[1] = 0xFFFFFFFFu;
  HAWK_UNROLL
  for (int k = 0; k < P; ++k) {
    const bool s0 = !((S.skip[0] >> k) & 1u), s1 = !((S.skip[1] >> k) & 1u);
    if (!(s0 || s1)) continue;
    const uint32_t a = funnel_r(cur.a, nxt.a, (uint32_t)k), c = funnel_r(cur.c, nxt.c, (uint32_t)k),
                   g = funnel_r(cur.g, nxt.g, (uint32_t)k), t = funnel_r(cur.t, nxt.t, (uint32_t)k);
    if (s0) m[0] &= (a & S.sel[0][k][0]) | (c & S.sel[0][k][1]) | (g & S.sel[0][k][2]) | (t & S.sel[0][k][3]);
    if (s1) m[1] &= (a & S.sel[1][k][0]) | (c & S.sel[1][k][1]) | (g & S.sel[1][k][2]) | (t & S.sel[1][k][3]);
  }
}

HAWK_HD void match_loop(const Planes& cur, const Planes& nxt, const PamSelect& S, int P, uint32_t m[2]) {
  m[0] = m[1] = 0xFFFFFFFFu;
  for (int k = 0; k < P; ++k) {
    const bool s0 = !((S.skip[0] >> k) & 1u), s1 = !((S.skip[1] >> k) & 1u);
    if (!(s0 || s1)) continue;
    const uint32_t a = funnel_r(cur.a, nxt.a, (uint32_t)k), c = funnel_r(cur.c, nxt.c, (uint32_t)k),
                   g = funnel_r(cur.g, nxt.g, (uint32_t)k), t = funnel_r(cur.t, nxt.t, (uint32_t)k);
    if (s0) m[0] &= (a & S.sel[0][k][0]) | (c & S.sel[0][k][1]) | (g & S.sel[0][k][2]) | (t & S.sel[0][k][3]);
    if (s1) m[1] &= (a & S.sel[1][k][0]) | (c & S.sel[1][k][1]) | (g & S.sel[1][k][2]) | (t & S.sel[1][k][3]);
  }
}

// the common PAM lengths get a fully unrolled matcher (P is uniform over the launch)
HAWK_HD void match_chunk2(const Planes& cur, const Planes& nxt, const PamSelect& S, int P, uint32_t m[2]) {
  switch (P) {
    case 2: match_fixed<2>(cur, nxt, S, m); break;
    case 3: match_fixed<3>(cur, nxt, S, m); break;
    case 4: match_fixed<4>(cur, nxt, S, m); break;
    case 5: match_fixed<5>(cur, nxt, S, m); break;
    case 6: match_fixed<6>(cur, nxt, S, m); break;
    default: match_loop(cur, nxt, S, P, m); break;
  }
}

// Per-strand geometry of search_guides.py:134-160, :395-420, :372-392.
// rp = right' (right XOR strand, :538). Window [pos + w0, pos + w1), core
// [pos + c0, pos + c0 + C), pivot = pos + c0.
struct StrandGeom {
  int32_t c0;   // core start relative to pos: 0 (right') or -G
  int32_t w0;   // window start relative to pos
  int32_t lo;   // smallest admissible pos (window inside the haplotype)
  int32_t hi_sub;  // admissible pos <= len - hi_sub
  int32_t stop_off;  // genomic stop = posmap[pos + stop_off] (:278-279)
};

HAWK_HD StrandGeom strand_geom(int G, int P, bool rp, bool unphased, int pad) {
  StrandGeom s;
  if (rp) {
    s.c0 = 0;
    s.w0 = -pad;
    s.lo = pad;                                   // pos - PAD >= 0
    s.hi_sub = G + P + pad + (unphased ? 1 : 0);  // pos + G + P + PAD <= len (< len when unphased)
    s.stop_off = G + P;
  } else {
    s.c0 = -G;
    s.w0 = -G - pad;
    s.lo = G + pad;      // pos - G - PAD >= 0 (is_pamhit_valid is the same bound)
    s.hi_sub = P + pad;  // pos + P + PAD <= len
    s.stop_off = P;
  }
  return s;
}

// Sliding OR over a 96-bit window held in three words (bit 0 = LSB of w0):
// afterwards bit x of the window = OR of the original bits [x, x + C), for every x with
// x + C <= 96. Log-doubling, 1 <= C <= 33 (all shifts are < 32).
HAWK_HD void dilate96_step(uint32_t& w0, uint32_t& w1, uint32_t& w2, uint32_t k) {
  w0 |= funnel_r(w0, w1, k);
  w1 |= funnel_r(w1, w2, k);
  w2 |= w2 >> k;
}
HAWK_HD void dilate96(uint32_t& w0, uint32_t& w1, uint32_t& w2, int C) {
  // after the steps taken so far every bit covers `cover` original bits; C is uniform
  // over the launch, so these are uniform branches
  if (C >= 2) dilate96_step(w0, w1, w2, 1);
  if (C >= 4) dilate96_step(w0, w1, w2, 2);
  if (C >= 8) dilate96_step(w0, w1, w2, 4);
  if (C >= 16) dilate96_step(w0, w1, w2, 8);
  if (C >= 32) dilate96_step(w0, w1, w2, 16);
  const int cover = C >= 32 ? 32 : C >= 16 ? 16 : C >= 8 ? 8 : C >= 4 ? 4 : C >= 2 ? 2 : 1;
  if (C > cover) dilate96_step(w0, w1, w2, (uint32_t)(C - cover));
}

// bit x of a bit-vector stored as 32-bit words, via accessor f(word_index)
template <class F>
HAWK_HD uint32_t bits32_at(F&& word, int64_t bit) {
  int64_t w = bit >> 5;  // arithmetic shift: floor for negatives
  uint32_t sh = (uint32_t)(bit & 31);
  uint32_t lo = word(w);
  if (sh == 0) return lo;
  return funnel_r(lo, word(w + 1), sh);
}

// search_guides.py:468-471 for 32 positions at once: bit i <=> some variant
// (lower-case) bit inside [p0 + i + c0, p0 + i + c0 + C). `vword(k)` returns
// case word k of the haplotype (0 outside it).
template <class F>
HAWK_HD uint32_t core_variant_mask(F&& vword, int64_t p0, int c0, int C) {
  // sliding OR by doubling over a 32-bit lane view: S_w(x) = OR v[x .. x+w)
  // evaluated for x = p0 + c0 + i, i in [0,32). Needs bits up to x + C - 1.
  // Work on an explicit window of words starting at the word holding p0 + c0.
  int64_t b0 = p0 + c0;
  int64_t w0 = b0 >> 5;
  uint32_t sh = (uint32_t)(b0 & 31);
  uint32_t out = 0;
  int nwords = (int)((sh + 31 + C + 31) >> 5);  // words covering [b0, b0 + 32 + C - 1)
  uint32_t prev = vword(w0);
  // out[i] = OR_{j<C} bit(b0 + i + j): walk the words once, OR-ing the C shifted views
  // that fall into each word pair. Equivalent to a dense loop over j but touches
  // each word pair once.
  int j = 0;
  for (int k = 0; k < nwords && j < C; ++k) {
    uint32_t next = vword(w0 + k + 1);
    // shifts s = sh + j - 32k must lie in [0, 32)
    for (; j < C; ++j) {
      int s = (int)sh + j - 32 * k;
      if (s >= 32) break;
      out |= funnel_r(prev, next, (uint32_t)s);
    }
    prev = next;
  }
  return out;
}

// haplotype.py:90-104,138-159 posmap, run-length encoded (hawkscan.h):
// evaluate posmap(i) for one haplotype's segment range [s0, s1).
HAWK_HD int32_t posmap_eval(const int32_t* seg_rel, const int32_t* seg_gen, const uint8_t* seg_step,
                            int64_t s0, int64_t s1, int32_t i) {
  // last segment with seg_rel <= i
  int64_t lo = s0, hi = s1;  // invariant: seg_rel[lo] <= i (seg_rel[s0] == 0)
  while (hi - lo > 1) {
    int64_t mid = (lo + hi) >> 1;
    if (seg_rel[mid] <= i) lo = mid; else hi = mid;
  }
  return seg_gen[lo] + (seg_step[lo] ? (i - seg_rel[lo]) : 0);
}

// extract nibble of slot `pos` of a haplotype whose first chunk is chunk0
HAWK_HD uint32_t nibble_at(const Planes* q, int64_t chunk0, int64_t pos) {
  const Planes& p = q[chunk0 + (pos >> 5)];
  uint32_t b = (uint32_t)(pos & 31);
  return ((p.a >> b) & 1u) | (((p.c >> b) & 1u) << 1) | (((p.g >> b) & 1u) << 2) |
         (((p.t >> b) & 1u) << 3);
}
HAWK_HD uint32_t lower_at(const uint32_t* v, int64_t chunk0, int64_t pos) {
  return (v[chunk0 + (pos >> 5)] >> (pos & 31)) & 1u;
}

// search_guides.py:356-369: upper-cased core equality == nibble equality over C slots
HAWK_HD bool cores_equal(const Planes* q, int64_t chunk0_a, int64_t pos_a, int64_t chunk0_b,
                         int64_t pos_b, int C) {
  for (int done = 0; done < C; done += 32) {
    const int n = C - done < 32 ? C - done : 32;
    const uint32_t keep = n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u);
    const int64_t ba = pos_a + done, bb = pos_b + done;
    const Planes a0 = q[chunk0_a + (ba >> 5)], a1 = q[chunk0_a + (ba >> 5) + 1];
    const Planes b0 = q[chunk0_b + (bb >> 5)], b1 = q[chunk0_b + (bb >> 5) + 1];
    const uint32_t sa = (uint32_t)(ba & 31), sb = (uint32_t)(bb & 31);
    const uint32_t diff = (funnel_r(a0.a, a1.a, sa) ^ funnel_r(b0.a, b1.a, sb)) |
                          (funnel_r(a0.c, a1.c, sa) ^ funnel_r(b0.c, b1.c, sb)) |
                          (funnel_r(a0.g, a1.g, sa) ^ funnel_r(b0.g, b1.g, sb)) |
                          (funnel_r(a0.t, a1.t, sa) ^ funnel_r(b0.t, b1.t, sb));
    if (diff & keep) return false;
  }
  return true;
}

}  // namespace hawk

// =============================================================================
// Views over one packed batch (device or host pointers) and the per-chunk scan.
// =============================================================================
#include "../../include/hawkscan.h"

namespace hawk {

#define HAWK_SEG_IDX_SHIFT 12

struct BatchView {
  const Planes* q;            // planes, one per chunk
  const uint32_t* v;          // case bits, one word per chunk
  const uint32_t* nz;         // summary of v: bit c of the plane = chunk c holds a variant base
  const int64_t* slot_off;    // n_hap + 1
  const int32_t* len;         // n_hap
  const int32_t* scan_start;  // n_hap (search_guides.py:49-84)
  const int32_t* scan_stop;   // n_hap
  const uint8_t* is_ref;      // n_hap (samples == "REF")
  int32_t n_hap;
  // coordinate maps (may be null for scan-only use)
  const int64_t* seg_off;
  const int32_t* seg_rel;
  const int32_t* seg_gen;
  const uint8_t* seg_step;
  // coarse index of the segments (may be null): seg_idx[h * seg_idx_stride + k] = last segment of
  // haplotype h (haplotype-local) that starts at or before relative index k << HAWK_SEG_IDX_SHIFT
  const int32_t* seg_idx;
  int32_t seg_idx_stride;
  // variant alleles (unphased only, may be null)
  const int64_t* va_off;
  const int32_t* va_idx;
  const int64_t* va_ent_off;
  const uint8_t* va_ref;
};

struct ScanConst {
  int32_t P, G, C;      // C = G + P (core length)
  int32_t right;        // --right
  int32_t unphased;     // HAWK_F_UNPHASED
  int32_t raw;          // 1: pam_search semantics (no in-range / REF-core filter)
  int32_t small;        // G <= 32 and C <= 33: a guide core reaches at most one case word either side
  int32_t back, ahead;  // case words before / after a chunk that can hold a core's variant base
  uint32_t prev_mask, next_mask;  // K.small: bits of the previous / next case word within reach
  uint8_t pat[2][HAWK_MAX_PAM];  // [0] forward PAM, [1] reverse complement
  PamSelect sel;                 // plane-select masks of pat
  StrandGeom geom[2];   // per strand (right' = right XOR strand)
};

HAWK_HD ScanConst make_scan_const(const hawk_params& p, int raw) {
  ScanConst k;
  k.P = p.pam_len;
  k.G = p.guide_len;
  k.C = p.pam_len + p.guide_len;
  k.right = p.right ? 1 : 0;
  k.unphased = (p.flags & HAWK_F_UNPHASED) ? 1 : 0;
  k.raw = raw;
  k.small = (k.G <= 32 && k.C <= 33) ? 1 : 0;
  k.back = (k.G + 31) >> 5;       // left core starts G bases before the PAM position
  k.ahead = (31 + k.C - 1) >> 5;  // right core ends C - 1 bases after it
  // chunk c holds a candidate position iff a variant bit lies in [32c - G, 32c + 31 + C - 1]
  k.prev_mask = k.G >= 32 ? 0xFFFFFFFFu : (k.G <= 0 ? 0u : ~0u << (32 - k.G));
  k.next_mask = (k.C - 1) >= 32 ? 0xFFFFFFFFu : ((1u << (k.C - 1)) - 1u);
  for (int i = 0; i < HAWK_MAX_PAM; ++i) {
    k.pat[0][i] = p.pam_fwd[i];
    k.pat[1][i] = p.pam_rc[i];
  }
  for (int s = 0; s < 2; ++s) {
    k.sel.skip[s] = 0;
    for (int i = 0; i < HAWK_MAX_PAM; ++i) {
      for (int b = 0; b < 4; ++b) k.sel.sel[s][i][b] = ((k.pat[s][i] >> b) & 1) ? 0xFFFFFFFFu : 0u;
      if (k.pat[s][i] == 15) k.sel.skip[s] |= 1u << i;
    }
  }
  for (int s = 0; s < 2; ++s)
    k.geom[s] = strand_geom(k.G, k.P, (k.right != 0) != (s == 1), k.unphased != 0, HAWK_GUIDESEQPAD);
  return k;
}

// per-haplotype scalars the scan needs, loaded once per span
struct HapScan {
  int64_t chunk0;   // first chunk of the haplotype
  int32_t len, a, b;
  int32_t nchunks;  // chunks holding its bases
  int32_t is_ref;
  int32_t lo[2], hi[2];  // admissible position interval per strand, [lo, hi)
  int32_t c_in_lo, c_in_hi;  // chunks [c_in_lo, c_in_hi) lie inside all three intervals
};

HAWK_HD HapScan load_hap_scan(const BatchView& B, const ScanConst& K, int32_t h) {
  HapScan H;
  H.chunk0 = B.slot_off[h] >> 5;
  H.len = B.len[h];
  H.a = B.scan_start[h];
  int32_t b = B.scan_stop[h];
  int32_t bmax = H.len - K.P + 1;  // a PAM must fit inside the haplotype
  H.b = b < bmax ? b : bmax;
  if (H.a < 0) H.a = 0;
  H.nchunks = (H.len + 31) >> 5;
  H.is_ref = B.is_ref[h];
  for (int s = 0; s < 2; ++s) {
    int32_t lo = H.a, hi = H.b;
    if (!K.raw) {
      int32_t glo = K.geom[s].lo, ghi = H.len - K.geom[s].hi_sub + 1;
      if (glo > lo) lo = glo;
      if (ghi < hi) hi = ghi;
    }
    H.lo[s] = lo;
    H.hi[s] = hi;
  }
  int32_t in_lo = H.a > H.lo[0] ? H.a : H.lo[0], in_hi = H.b < H.hi[0] ? H.b : H.hi[0];
  if (H.lo[1] > in_lo) in_lo = H.lo[1];
  if (H.hi[1] < in_hi) in_hi = H.hi[1];
  H.c_in_lo = in_lo <= 0 ? 0 : (in_lo + 31) >> 5;
  H.c_in_hi = in_hi <= 0 ? 0 : in_hi >> 5;
  return H;
}

// K.small form of scan_chunk below with the three case words of chunks c - 1, c, c + 1
// handed in (the kernel queues them with the candidate); use_v = false for REF / pam_search
// mode, where the case words play no role.
// ... with the planes of chunks c and c + 1 handed in as well (the fused kernel loads them beside
// the haplotype's scan geometry instead of after it)
HAWK_HD void scan_chunk_small_pre(const ScanConst& K, const HapScan& H, int32_t c, uint32_t w0, uint32_t w1,
                                  uint32_t w2, bool use_v, const Planes& cur, const Planes& nxt, uint32_t out[2],
                                  uint32_t raw[2]) {
  const int32_t p0 = c << 5;
  out[0] = out[1] = raw[0] = raw[1] = 0;
  uint32_t inscan = 0xFFFFFFFFu, cand[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
  if (c < H.c_in_lo || c >= H.c_in_hi) {  // boundary chunk of the scan / window intervals
    inscan = interval_mask(H.a, H.b, p0);
    if (!inscan) return;
    cand[0] = interval_mask(H.lo[0], H.hi[0], p0);
    cand[1] = interval_mask(H.lo[1], H.hi[1], p0);
  }
  if (use_v) {
    dilate96(w0, w1, w2, K.C);
    HAWK_UNROLL
    for (int s = 0; s < 2; ++s) {
      const uint32_t off = (uint32_t)(32 + K.geom[s].c0);
      cand[s] &= (off == 32u) ? w1 : funnel_r(w0, w1, off);
    }
    if (!(cand[0] | cand[1])) return;
  }
  uint32_t m[2];
  match_chunk2(cur, nxt, K.sel, K.P, m);
  raw[0] = m[0] & inscan;
  raw[1] = m[1] & inscan;
  out[0] = m[0] & cand[0];
  out[1] = m[1] & cand[1];
}

HAWK_HD void scan_chunk_small(const BatchView& B, const ScanConst& K, const HapScan& H, int32_t c, uint32_t w0,
                              uint32_t w1, uint32_t w2, bool use_v, uint32_t out[2], uint32_t raw[2]) {
  const int32_t p0 = c << 5;
  out[0] = out[1] = raw[0] = raw[1] = 0;
  uint32_t inscan = 0xFFFFFFFFu, cand[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
  if (c < H.c_in_lo || c >= H.c_in_hi) {  // boundary chunk of the scan / window intervals
    inscan = interval_mask(H.a, H.b, p0);
    if (!inscan) return;
    cand[0] = interval_mask(H.lo[0], H.hi[0], p0);
    cand[1] = interval_mask(H.lo[1], H.hi[1], p0);
  }
  if (use_v) {
    dilate96(w0, w1, w2, K.C);
    HAWK_UNROLL
    for (int s = 0; s < 2; ++s) {
      // core of position p0 + i starts at window bit 32 + c0 + i (c0 = 0 or -G, G <= 32)
      const uint32_t off = (uint32_t)(32 + K.geom[s].c0);
      cand[s] &= (off == 32u) ? w1 : funnel_r(w0, w1, off);
    }
    if (!(cand[0] | cand[1])) return;
  }
  const Planes cur = B.q[H.chunk0 + c], nxt = B.q[H.chunk0 + c + 1];
  uint32_t m[2];
  match_chunk2(cur, nxt, K.sel, K.P, m);
  raw[0] = m[0] & inscan;
  raw[1] = m[1] & inscan;
  out[0] = m[0] & cand[0];
  out[1] = m[1] & cand[1];
}

// The whole per-position work of pam_search + the fused filters for the 32
// positions of chunk `c` (haplotype-relative) : out[s] = surviving hit bits of
// strand s, raw[s] = PAM matches inside the scan interval before the filters.
// `vword(w)` returns case word w of the haplotype (0 outside it); the kernel serves it
// from the shared-memory tile, the host check from the plane in memory.
template <class VW>
HAWK_HD void scan_chunk(const BatchView& B, const ScanConst& K, const HapScan& H, int64_t c,
                        VW&& vword, uint32_t out[2], uint32_t raw[2]) {
  const int32_t p0 = (int32_t)(c << 5);
  out[0] = out[1] = raw[0] = raw[1] = 0;
  uint32_t inscan = 0xFFFFFFFFu, cand[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
  if (c < H.c_in_lo || c >= H.c_in_hi) {  // boundary chunk of the scan / window intervals
    inscan = interval_mask(H.a, H.b, p0);
    if (!inscan) return;
    cand[0] = interval_mask(H.lo[0], H.hi[0], p0);
    cand[1] = interval_mask(H.lo[1], H.hi[1], p0);
  }
  if (!K.raw && !H.is_ref) {
    // search_guides.py:468-471: a non-REF hit survives only if its core holds a
    // variant base; skip the plane loads when no variant bit is in reach.
    if (K.small) {
      uint32_t w0 = vword(c - 1), w1 = vword(c), w2 = vword(c + 1);
      if (!(w0 | w1 | w2)) return;
      dilate96(w0, w1, w2, K.C);
      HAWK_UNROLL
      for (int s = 0; s < 2; ++s) {
        // core of position p0 + i starts at window bit 32 + c0 + i (c0 = 0 or -G, G <= 32)
        uint32_t off = (uint32_t)(32 + K.geom[s].c0);
        cand[s] &= (off == 32u) ? w1 : funnel_r(w0, w1, off);
      }
    } else {
      for (int s = 0; s < 2; ++s)
        if (cand[s]) cand[s] &= core_variant_mask(vword, p0, K.geom[s].c0, K.C);
    }
    if (!(cand[0] | cand[1])) return;
  }
  Planes cur = B.q[H.chunk0 + c], nxt = B.q[H.chunk0 + c + 1];
  uint32_t m[2];
  match_chunk2(cur, nxt, K.sel, K.P, m);
  raw[0] = m[0] & inscan;
  raw[1] = m[1] & inscan;
  out[0] = m[0] & cand[0];
  out[1] = m[1] & cand[1];
}

// same, case words read straight from the batch's plane
HAWK_HD void scan_chunk(const BatchView& B, const ScanConst& K, const HapScan& H, int64_t c,
                        uint32_t out[2], uint32_t raw[2]) {
  auto vword = [&](int64_t w) -> uint32_t {
    return (w < 0 || w >= H.nchunks) ? 0u : B.v[H.chunk0 + w];
  };
  scan_chunk(B, K, H, c, vword, out, raw);
}

// ---- per-hit row geometry -------------------------------------------------------
struct RowCoords {
  int32_t pivot;  // first core index (search_guides.py:301)
  int32_t start, stop;  // adjust_guide_position, :260-280
};

// entry k of haplotype h's coarse segment index: the last segment (haplotype-local) that starts at or
// before relative index k << HAWK_SEG_IDX_SHIFT
HAWK_HD int32_t seg_index_entry(const int64_t* seg_off, const int32_t* seg_rel, int32_t h, int32_t k) {
  const int64_t s0 = seg_off[h];
  const int32_t n = (int32_t)(seg_off[h + 1] - s0);
  const int32_t* rel = seg_rel + s0;
  const int64_t target = (int64_t)k << HAWK_SEG_IDX_SHIFT;
  int32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (rel[mid] <= target) lo = mid; else hi = mid;
  }
  return lo;
}

HAWK_HD RowCoords row_coords(const BatchView& B, const ScanConst& K, int32_t h, int32_t pos, int s) {
  RowCoords r;
  const StrandGeom& g = K.geom[s];
  r.pivot = pos + g.c0;
  // the haplotype's own segment arrays, indexed with 32-bit offsets (the search was a third of
  // rows_fast's instructions when it ran on 64-bit global indices)
  const int64_t s0 = B.seg_off[h];
  const int32_t n = (int32_t)(B.seg_off[h + 1] - s0);
  const int32_t* rel = B.seg_rel + s0;
  const int32_t* gen = B.seg_gen + s0;
  const uint8_t* step = B.seg_step + s0;
  // last segment with rel <= pivot, then walk forward to the one holding the stop index
  // (a guide spans G + P bases: almost always the same segment or the next)
  int32_t lo = 0, hi = n;
  if (B.seg_idx) {  // one lookup narrows the search to the segments of the pivot's 4 kb bucket
    int32_t k = r.pivot < 0 ? 0 : r.pivot >> HAWK_SEG_IDX_SHIFT;
    if (k > B.seg_idx_stride - 2) k = B.seg_idx_stride - 2;
    const int32_t* ix = B.seg_idx + (int64_t)h * B.seg_idx_stride + k;
    lo = ix[0];
    hi = ix[1] + 1;
  }
  while (hi - lo > 1) {
    const int32_t mid = (lo + hi) >> 1;
    if (rel[mid] <= r.pivot) lo = mid; else hi = mid;
  }
  r.start = gen[lo] + (step[lo] ? (r.pivot - rel[lo]) : 0);
  const int32_t j = pos + g.stop_off;
  while (lo + 1 < n && rel[lo + 1] <= j) ++lo;
  r.stop = gen[lo] + (step[lo] ? (j - rel[lo]) : 0);
  return r;
}

// Find the REF row (records of haplotype `ref_h`, ascending pos, in
// recs[lo, hi)) whose genomic start equals `start`; returns its index or -1.
// The start of a REF row is monotone in pos, so binary search applies.
HAWK_HD int64_t find_ref_partner(const BatchView& B, const ScanConst& K, const uint64_t* recs,
                                 int64_t lo, int64_t hi, int32_t ref_h, int s, int32_t start) {
  int64_t s0 = B.seg_off[ref_h], s1 = B.seg_off[ref_h + 1];
  int c0 = K.geom[s].c0;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int32_t pos = (int32_t)(recs[mid] & 0xFFFFFFFFu);
    int32_t st = posmap_eval(B.seg_rel, B.seg_gen, B.seg_step, s0, s1, pos + c0);
    if (st < start) lo = mid + 1; else hi = mid;
  }
  return lo;  // caller checks lo < end and start equality
}

// ---- N2: post-search pure functions on one guide row (annotation.py:27-51, 197-281, 513-541) ----
// Variant table of a batch: haplotype h carries variants [var_off[h], var_off[h+1]), sorted by
// position, in the reference's normalised form (variant.py:456-486): var_pos + pos_base = genomic
// coordinate of the anchor, REF / ALT allele lengths, ALT text = alt_pool[var_altoff .. + altlen).
struct VariantView {
  const int64_t* var_off;
  const int32_t* var_pos;
  const int32_t* var_reflen;
  const int32_t* var_altlen;
  const int64_t* var_altoff;
  const uint8_t* alt_pool;
  int32_t pos_base;
};

HAWK_HD bool ascii_is_upper(uint8_t c) { return c >= 'A' && c <= 'Z'; }
HAWK_HD uint8_t ascii_to_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }

// complement of one IUPAC letter, case kept (utils.py:46-79); other bytes unchanged
HAWK_HD uint8_t rc_char(uint8_t c) {
  const uint8_t u = ascii_to_upper(c);
  uint8_t o;
  switch (u) {
    case 'A': o = 'T'; break;
    case 'C': o = 'G'; break;
    case 'G': o = 'C'; break;
    case 'T': o = 'A'; break;
    case 'U': o = 'A'; break;
    case 'R': o = 'Y'; break;
    case 'Y': o = 'R'; break;
    case 'M': o = 'K'; break;
    case 'K': o = 'M'; break;
    case 'H': o = 'D'; break;
    case 'D': o = 'H'; break;
    case 'B': o = 'V'; break;
    case 'V': o = 'B'; break;
    default: o = u; break;  // N, S, W
  }
  return (uint8_t)(o | (c & 0x20));
}

// byte i of the text reverse_guides leaves in a guide of strand s (window of W characters)
HAWK_HD uint8_t annot_text_byte(const uint8_t* src, int W, int s, int i) {
  if (i >= W) return 0;
  return s ? rc_char(src[W - 1 - i]) : src[i];
}

// gc_content: G+C+S and A+C+G+T+S+W counts of the guide without its PAM (forward text; the
// counts are invariant under reverse complement)
HAWK_HD void annot_gc_counts(const ScanConst& K, int s, const uint8_t* src, int32_t* num_out, int32_t* den_out) {
  const bool rp = (K.right != 0) != (s == 1);  // right' = right XOR strand (search_guides.py:538)
  const int g0 = HAWK_GUIDESEQPAD + (rp ? K.P : 0);
  int num = 0, den = 0;
  for (int i = 0; i < K.G; ++i) {
    const uint8_t u = ascii_to_upper(src[g0 + i]);
    const int gc = (u == 'G') | (u == 'C') | (u == 'S');
    num += gc;
    den += gc | (u == 'A') | (u == 'T') | (u == 'W');
  }
  *num_out = num;
  *den_out = den;
}

// ---- N4: CFDon of one guide row against the REF guide of its (start, strand) key ----------------
// scores/cfdscore/cfdscore.py:53-95 (compute_cfd) as called by scores/crisprhawk_scores.py:65-87
// (cfdon): score = 1.0, times mm[i][wildtype[i]][guide[i]] for every i < min(G, 20) where the
// upper-cased letters differ (key "r<W>:d<revcomp(S)>,<i+1>" of the reference's table, T read as
// U), times pam2[last two PAM letters], multiplied in that order in double precision -- the same
// IEEE operations as the Python loop. Both texts are the ones reverse_guides leaves (strand 1:
// reverse complement), guide and PAM laid out by `right`. mm: 20 x 4 x 4 (A, C, G, T/U), pam2:
// 4 x 4; a NaN entry stands for a key the reference's dict does not hold. Returns false where the
// reference raises KeyError (a mismatch on a letter other than A, C, G, T; a PAM shorter than two
// letters or with such a letter; a missing table entry).
HAWK_HD int acgt_code(uint8_t c) {
  switch (ascii_to_upper(c)) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': case 'U': return 3;
    default: return -1;
  }
}

HAWK_HD bool cfdon_row(const uint8_t* ref_text, const uint8_t* row_text, int W, int G, int P, int right, int s,
                       const double* mm, const double* pam2, double* out) {
  const int g0 = HAWK_GUIDESEQPAD + (right ? P : 0), p0 = HAWK_GUIDESEQPAD + (right ? 0 : G);
  double score = 1.0;
  const int n = G < 20 ? G : 20;
  for (int i = 0; i < n; ++i) {
    const uint8_t w = ascii_to_upper(annot_text_byte(ref_text, W, s, g0 + i));
    const uint8_t g = ascii_to_upper(annot_text_byte(row_text, W, s, g0 + i));
    if (w == g) continue;
    const int cw = acgt_code(w), cg = acgt_code(g);
    if (cw < 0 || cg < 0) return false;
    const double f = mm[(i * 4 + cw) * 4 + cg];
    if (f != f) return false;
    score *= f;
  }
  if (P < 2) return false;
  const int c0 = acgt_code(annot_text_byte(row_text, W, s, p0 + P - 2));
  const int c1 = acgt_code(annot_text_byte(row_text, W, s, p0 + P - 1));
  if (c0 < 0 || c1 < 0) return false;
  const double f = pam2[c0 * 4 + c1];
  if (f != f) return false;
  *out = score * f;
  return true;
}

// ---- N4 (second half): the learned scorers' input window of one guide row ----------------------
// scoring.py:50-84 (_extract_guide_sequences / _extract_guide_sequences_sgdesigner):
// sequence[(PAD - lead) : (-PAD + 3)].upper() of the text reverse_guides leaves (strand 1: reverse
// complement), lead = 4 (Azimuth, RS3, DeepCpf1, CRISPRon) or 0 (sgDesigner); G + P + lead + 3
// letters. Byte j of that string:
HAWK_HD int feature_len(int W, int lead) { return W - 2 * HAWK_GUIDESEQPAD + lead + 3; }
HAWK_HD uint8_t feature_byte(const uint8_t* src, int W, int s, int lead, int j) {
  return ascii_to_upper(annot_text_byte(src, W, s, HAWK_GUIDESEQPAD - lead + j));
}
// scores/deepCpf1/seqdeepcpf1.py:19, 71-92 (preprocess): channel of a letter in the one-hot
// tensor, A 0, C 1, G 2, T 3; -1 where NTENCODING raises KeyError (U is not a key there)
HAWK_HD int onehot_channel(uint8_t upper) {
  switch (upper) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': return 3;
    default: return -1;
  }
}

// polish_guide_variants (annotation.py:246-281) for one row: walks the core's G + P bases,
// genomic coordinate through the run-length posmap (segment pointer advanced incrementally),
// variant at that coordinate by a monotone walk of the haplotype's sorted table, then
// _check_insertion / _check_snv (:197-243) on the window text. emit(j) receives the
// haplotype-local index of every variant kept (a set: once each); returns their number.
// *would_assert is set when the reference's _find_insertion_stop assert would fire.
template <class Emit>
HAWK_HD uint32_t annot_row_variants(const BatchView& B, const ScanConst& K, const VariantView& V, int32_t h, int s,
                                    int32_t pos, int32_t stop, const uint8_t* core, Emit&& emit, bool* would_assert) {
  const int32_t pivot = pos + K.geom[s].c0;
  const int C = K.C;
  const int64_t v0 = V.var_off[h], v1 = V.var_off[h + 1];
  uint32_t found = 0;
  if (v1 <= v0) return 0;
  // segment holding the core's first base
  const int64_t s0 = B.seg_off[h], s1 = B.seg_off[h + 1];
  int64_t k = s0, hi = s1;
  while (hi - k > 1) {
    const int64_t mid = (k + hi) >> 1;
    if (B.seg_rel[mid] <= pivot) k = mid; else hi = mid;
  }
  int32_t p = B.seg_gen[k] + (B.seg_step[k] ? (pivot - B.seg_rel[k]) : 0);
  // first variant at or after the core's first coordinate
  int64_t j = v0, jh = v1;
  while (j < jh) {
    const int64_t mid = (j + jh) >> 1;
    if (V.var_pos[mid] + V.pos_base < p) j = mid + 1; else jh = mid;
  }
  int64_t last = -1;
  for (int i = 0; i < C && j < v1; ++i) {
    const int32_t idx = pivot + i;
    while (k + 1 < s1 && B.seg_rel[k + 1] <= idx) ++k;
    p = B.seg_gen[k] + (B.seg_step[k] ? (idx - B.seg_rel[k]) : 0);
    while (j < v1 && V.var_pos[j] + V.pos_base < p) ++j;
    int offset = 0;  // annotation.py:264: reset per base, carried over the variants of one base
    for (int64_t jj = j; jj < v1 && V.var_pos[jj] + V.pos_base == p; ++jj) {
      const int32_t rl = V.var_reflen[jj], al = V.var_altlen[jj];
      const uint8_t* alt = V.alt_pool + V.var_altoff[jj];
      const bool is_snv = rl == al;
      if (!is_snv) offset = rl < al ? al - rl : 0;
      const int seglen = (i + offset + 1 <= C ? offset + 1 : C - i);
      const uint8_t* seg = core + i;
      bool ok = false;
      if (!is_snv) {  // _check_insertion (:197-226)
        if (i == 0) {
          int up = -1;  // _find_insertion_stop: first upper-case character, 0 when none
          for (int t = 0; t < seglen; ++t)
            if (ascii_is_upper(seg[t])) { up = t; break; }
          if (up == 0) *would_assert = true;  // the reference asserts here
          const int kk = up < 0 ? 0 : up;
          bool e = kk <= al;
          for (int t = 0; e && t < kk; ++t) e = alt[al - kk + t] == ascii_to_upper(seg[t]);
          ok = e;
        }
        if (!ok && p == stop) {
          bool e = seglen <= al;
          for (int t = 0; e && t < seglen; ++t) e = alt[t] == ascii_to_upper(seg[t]);
          ok = e;
        }
      }
      if (!ok) {  // _check_snv (:229-243): all lower-case and equal to the ALT allele
        bool e = seglen == al;
        for (int t = 0; e && t < seglen; ++t) e = !ascii_is_upper(seg[t]) && alt[t] == ascii_to_upper(seg[t]);
        ok = e;
      }
      if (ok && jj != last) {  // a set: the bases of one insertion share the anchor's coordinate
        last = jj;
        emit((int32_t)(jj - v0));
        ++found;
      }
    }
  }
  return found;
}

// ---- unphased resolution (search_guides.py:175-257) -------------------------------
// Column descriptor of one window position.
struct Column {
  uint32_t nib;      // IUPAC nibble of the haplotype base
  uint32_t lower;    // haplotype base is lower-case
  uint32_t allowed;  // bases (one-hot subset of nib) allowed: nib & PAM pattern on PAM columns
  int64_t e0;        // first allele entry (ambiguous columns)
  int32_t m;         // number of allele entries (0 => column is a single concrete base)
};

// returns false when an ambiguity code has no variant_alleles entry (KeyError)
HAWK_HD bool load_column(const BatchView& B, int32_t h, int64_t chunk0, int32_t idx,
                         uint32_t pam_code /* 0 = not a PAM column */, Column& c) {
  c.nib = nibble_at(B.q, chunk0, idx);
  c.lower = lower_at(B.v, chunk0, idx);
  c.m = 0;
  c.e0 = 0;
  c.allowed = pam_code ? (c.nib & pam_code) : c.nib;
  if (popc32(c.nib) <= 1) return true;
  // binary search the haplotype's allele sites for idx
  int64_t lo = B.va_off[h], hi = B.va_off[h + 1];
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    if (B.va_idx[mid] < idx) lo = mid + 1; else hi = mid;
  }
  if (lo >= B.va_off[h + 1] || B.va_idx[lo] != idx) return false;
  c.e0 = B.va_ent_off[lo];
  c.m = (int32_t)(B.va_ent_off[lo + 1] - c.e0);
  return true;
}

// number of candidate characters of a column that survive the PAM filter
HAWK_HD uint32_t column_count(const Column& c) {
  if (c.m == 0) return c.allowed ? 1u : 0u;
  return (uint32_t)popc32(c.allowed) * (uint32_t)c.m;
}

// t-th surviving candidate, in the reference's order: bases ascending A,C,G,T
// (utils.py:82-98), then allele entries (search_guides.py:207-213); upper-case
// iff the base equals the entry's REF allele.
HAWK_HD char column_char(const BatchView& B, const Column& c, uint32_t t) {
  if (c.m == 0) {
    char ch = nibble_letter(c.nib);
    return c.lower ? (char)(ch + 32) : ch;
  }
  uint32_t bi = t / (uint32_t)c.m, e = t % (uint32_t)c.m;
  uint32_t base = 0, allowed = c.allowed;
  for (uint32_t k = 0; k <= bi; ++k) {  // bi-th set bit of `allowed`
    base = allowed & (~allowed + 1u);
    allowed &= allowed - 1u;
  }
  char ch = nibble_letter(base);
  return (B.va_ref[c.e0 + e] == base) ? ch : (char)(ch + 32);
}

}  // namespace hawk
