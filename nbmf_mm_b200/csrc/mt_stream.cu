// Host-side MT19937 with arbitrary jump-ahead: the init stream of the reference solver, produced for ONE row shard.
//
// The reference seeds NumPy's legacy global generator and draws W_init (m x k) and then H_init (k x n) from it
// (_solver.py:102-103,126-129).  A row-sharded fit needs rows [r0, r1) of W_init and all of H_init on every rank; drawing
// the whole of W_init on every rank just to reach its own rows (and H_init behind them) costs 0.25 s at 10^6 x 32 and does
// not shrink with the number of GPUs.  MT19937 is a linear recurrence over GF(2), so the state J outputs ahead is
// g(F) s with g(x) = x^J mod phi(x), phi the characteristic polynomial of the one-word transition F (Haramoto,
// Matsumoto, Nishimura, Panneton, L'Ecuyer 2008).  phi is found once per process by Berlekamp-Massey on an output bit
// sequence, g by square-and-multiply-by-x (cached per J), g(F) s by Horner's rule.  The doubles that come out are
// bit-identical to numpy.random.RandomState(seed).uniform(lo, hi, ...) at the same stream position
// (tests/test_mt_stream.py checks that against NumPy on the CPU).
#include <stdint.h>
#include <string.h>

#include <map>
#include <mutex>
#include <vector>

#include "../../include/nbmf_b200.h"

namespace {

constexpr int N = 624, M = 397, DEG = 19937, PW = (DEG + 64) / 64;   // PW words hold degrees 0..DEG
constexpr uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;

struct Mt {
  uint32_t mt[N];
  int pos;   // next word to temper; N = regenerate first (NumPy's RK_STATE_LEN convention)
};

void mt_seed(Mt& s, uint32_t seed) {          // init_genrand == numpy mt19937_seed
  s.mt[0] = seed;
  for (int i = 1; i < N; ++i) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
  s.pos = N;
}
void mt_regen(Mt& s) {
  uint32_t* mt = s.mt;
  int kk = 0;
  for (; kk < N - M; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; kk < N - 1; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  const uint32_t y = (mt[N - 1] & UPPER) | (mt[0] & LOWER);
  mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  s.pos = 0;
}
inline uint32_t mt_next32(Mt& s) {
  if (s.pos >= N) mt_regen(s);
  uint32_t y = s.mt[s.pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
inline double mt_next_double(Mt& s) {         // genrand_res53 == numpy mt19937_next_double
  const uint32_t a = mt_next32(s) >> 5, b = mt_next32(s) >> 6;
  return (a * 67108864.0 + b) / 9007199254740992.0;
}

// ---- the recurrence in "sliding window" form for the jump: the window x[0..623] with x[623] the newest word.
// One step appends x_new = x[M] ^ twist(x[0], x[1]) and drops x[0].  A circular buffer keeps it O(1).
struct Win {
  uint32_t x[N];
  int head;   // index of the oldest word
  inline uint32_t at(int i) const { int j = head + i; return x[j >= N ? j - N : j]; }
  inline void step() {
    const uint32_t y = (at(0) & UPPER) | (at(1) & LOWER);
    const uint32_t v = at(M) ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
    x[head] = v;                  // the oldest slot becomes the newest word
    head = head + 1 == N ? 0 : head + 1;
  }
};

// ---- GF(2)[x] helpers on bit-packed polynomials
using Poly = std::vector<uint64_t>;
inline int bit(const Poly& p, int i) { return (int)((p[(size_t)i >> 6] >> (i & 63)) & 1u); }

std::once_flag g_phi_once;
Poly g_phi;                                  // PW words, degree DEG

// Berlekamp-Massey over GF(2) on 2 * DEG bits of the output sequence "bit 0 of every new word".
void compute_phi() {
  const int L2 = 2 * DEG + 2;
  std::vector<uint64_t> seq((size_t)(L2 + 63) / 64, 0);
  Win w;
  Mt s;
  mt_seed(s, 5489u);
  memcpy(w.x, s.mt, sizeof(w.x));
  w.head = 0;
  for (int i = 0; i < L2; ++i) {
    w.step();
    const int newest = w.head == 0 ? N - 1 : w.head - 1;
    if (w.x[newest] & 1u) seq[(size_t)i >> 6] |= 1ull << (i & 63);
  }
  // C(x): connection polynomial, B(x): previous; reversed-sequence trick for word-parallel discrepancy:
  // d = sum_{i=0..L} c_i s_{n-i}.  Keep the sequence reversed in a sliding bit window to AND word by word.
  const int W = (DEG + 2 + 63) / 64 + 1;
  std::vector<uint64_t> C((size_t)W, 0), B((size_t)W, 0), T((size_t)W), R((size_t)W, 0);   // R: bit i = s_{n-i}
  C[0] = B[0] = 1;
  int L = 0, m = 1;
  for (int n = 0; n < L2; ++n) {
    // shift R left by one and insert s_n at bit 0
    for (int i = W - 1; i > 0; --i) R[(size_t)i] = (R[(size_t)i] << 1) | (R[(size_t)i - 1] >> 63);
    R[0] = (R[0] << 1) | ((seq[(size_t)n >> 6] >> (n & 63)) & 1u);
    uint64_t acc = 0;
    const int lw = L / 64 + 1;
    for (int i = 0; i < lw && i < W; ++i) acc ^= C[(size_t)i] & R[(size_t)i];
    if (__builtin_parityll(acc)) {
      T = C;
      // C ^= B << m
      const int ws = m >> 6, bs = m & 63;
      for (int i = W - 1; i >= ws; --i) {
        uint64_t v = B[(size_t)(i - ws)] << bs;
        if (bs && i - ws - 1 >= 0) v |= B[(size_t)(i - ws - 1)] >> (64 - bs);
        C[(size_t)i] ^= v;
      }
      if (2 * L <= n) { L = n + 1 - L; B = T; m = 1; } else { ++m; }
    } else {
      ++m;
    }
  }
  // C(x) = 1 + c_1 x + ... + c_L x^L with s_n = sum c_i s_{n-i}; the characteristic polynomial is its reciprocal:
  // phi(x) = x^L C(1/x), coefficient of x^(L-i) is c_i.
  g_phi.assign((size_t)PW, 0);
  if (L != DEG) return;                      // callers check the degree
  for (int i = 0; i <= L; ++i)
    if ((C[(size_t)i >> 6] >> (i & 63)) & 1u) g_phi[(size_t)(L - i) >> 6] |= 1ull << ((L - i) & 63);
}

// r(x) = x^J mod phi(x), bit-packed (PW words).  Left-to-right binary exponentiation: square, then times x.
std::mutex g_cache_mu;
std::map<uint64_t, Poly> g_jump_cache;

void reduce_top(std::vector<uint64_t>& a, int top_deg) {       // a mod phi, a has degree <= top_deg
  const Poly& phi = g_phi;
  for (int d = top_deg; d >= DEG; --d) {
    if (!((a[(size_t)d >> 6] >> (d & 63)) & 1u)) continue;
    const int sh = d - DEG, ws = sh >> 6, bs = sh & 63;
    for (int i = 0; i < PW; ++i) {
      a[(size_t)(i + ws)] ^= phi[(size_t)i] << bs;
      if (bs) a[(size_t)(i + ws + 1)] ^= phi[(size_t)i] >> (64 - bs);
    }
  }
}
inline uint64_t spread32(uint32_t v) {        // interleave zeros: squaring in GF(2)[x]
  uint64_t x = v;
  x = (x | (x << 16)) & 0x0000ffff0000ffffull;
  x = (x | (x << 8)) & 0x00ff00ff00ff00ffull;
  x = (x | (x << 4)) & 0x0f0f0f0f0f0f0f0full;
  x = (x | (x << 2)) & 0x3333333333333333ull;
  x = (x | (x << 1)) & 0x5555555555555555ull;
  return x;
}
const Poly& jump_poly(uint64_t J) {
  std::lock_guard<std::mutex> lk(g_cache_mu);
  auto it = g_jump_cache.find(J);
  if (it != g_jump_cache.end()) return it->second;
  std::vector<uint64_t> r((size_t)2 * PW + 2, 0), sq((size_t)2 * PW + 2);
  r[0] = 1;                                   // x^0
  int top = 63;
  while (top > 0 && !((J >> top) & 1ull)) --top;
  for (int b = top; b >= 0; --b) {
    // square
    std::fill(sq.begin(), sq.end(), 0);
    for (int i = 0; i < PW; ++i) {
      sq[(size_t)2 * i] = spread32((uint32_t)r[(size_t)i]);
      sq[(size_t)2 * i + 1] = spread32((uint32_t)(r[(size_t)i] >> 32));
    }
    reduce_top(sq, 2 * (DEG - 1));
    if ((J >> b) & 1ull) {                    // times x
      for (int i = PW; i > 0; --i) sq[(size_t)i] = (sq[(size_t)i] << 1) | (sq[(size_t)i - 1] >> 63);
      sq[0] <<= 1;
      reduce_top(sq, DEG);
    }
    std::copy(sq.begin(), sq.begin() + PW + 1, r.begin());
    std::fill(r.begin() + PW + 1, r.end(), 0);
  }
  Poly out(r.begin(), r.begin() + PW);
  if (g_jump_cache.size() > 64) g_jump_cache.clear();
  return g_jump_cache.emplace(J, std::move(out)).first->second;
}

// Advance a freshly regenerated-or-not generator by J 32-bit outputs.
// The stream of raw words is x_0, x_1, ... with (mt[0..623], pos): the words still to be tempered are mt[pos..623]
// followed by the recurrence.  Normalise to a window whose oldest word is the next output, jump the window, and
// rebuild (mt, pos = 0 .. ) from it.
void mt_jump(Mt& s, uint64_t J) {
  if (J == 0) return;
  if (J < 4096) {                             // short hops: just step
    for (uint64_t i = 0; i < J; ++i) (void)mt_next32(s);
    return;
  }
  std::call_once(g_phi_once, compute_phi);
  // window W0 = the 624 words starting at the next output.  With pos = N (fresh seed) the generator regenerates first,
  // so the next outputs are the words AFTER mt[0..623]: make the state explicit by regenerating now.
  if (s.pos >= N) mt_regen(s);
  Win w;
  // next output is mt[pos]; the window of 624 consecutive raw words starting there = mt[pos..623] then new words.
  // Build it by stepping a window that starts at mt[0..623] (oldest = mt[0]) forward by pos words.
  memcpy(w.x, s.mt, sizeof(w.x));
  w.head = 0;
  for (int i = 0; i < s.pos; ++i) w.step();
  // NOTE on the unused bits: the recurrence never reads the low 31 bits of the oldest word, and the words the window
  // holds are genuine outputs of the recurrence, so linear combinations of shifted windows are windows of the stream.
  const Poly& g = jump_poly(J);
  // Horner: acc = sum_i g_i F^i(w)  evaluated as (((g_top F + g_{top-1}) F + ...) F + g_0) applied to w
  int top = DEG - 1;
  while (top > 0 && !bit(g, top)) --top;
  Win acc;
  if (bit(g, top)) { acc = w; } else { memset(acc.x, 0, sizeof(acc.x)); acc.head = 0; }
  for (int i = top - 1; i >= 0; --i) {
    acc.step();
    if (bit(g, i))
      for (int j = 0; j < N; ++j) {           // acc += w, aligned oldest to oldest
        int a = acc.head + j; if (a >= N) a -= N;
        int b = w.head + j; if (b >= N) b -= N;
        acc.x[a] ^= w.x[b];
      }
  }
  for (int j = 0; j < N; ++j) s.mt[j] = acc.at(j);
  s.pos = 0;
}

}  // namespace

// out[i] = lo + (hi - lo) * u_i for the `count` doubles that follow `skip` doubles of numpy.random.RandomState(seed)
// (legacy MT19937, two 32-bit outputs per double).  state_out (nullable, 625 uint32): mt[624] + pos afterwards.
extern "C" int nbmf_mt19937_uniform(uint32_t seed, uint64_t skip, uint64_t count, double lo, double hi, double* out,
                                    uint32_t* state_out) {
  if (count && !out) return NBMF_ERR_ARG;
  Mt s;
  mt_seed(s, seed);
  mt_jump(s, 2 * skip);
  if (skip >= 2048 && g_phi.empty()) return NBMF_ERR_UNSUPPORTED;
  if (skip >= 2048 && !(g_phi[(size_t)DEG >> 6] >> (DEG & 63) & 1u)) return NBMF_ERR_UNSUPPORTED;   // Berlekamp-Massey failed
  const double scale = hi - lo;
  for (uint64_t i = 0; i < count; ++i) out[i] = lo + scale * mt_next_double(s);
  if (state_out) {
    memcpy(state_out, s.mt, sizeof(s.mt));
    state_out[N] = (uint32_t)s.pos;
  }
  return NBMF_OK;
}
