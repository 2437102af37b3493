// rau_layout.cuh -- flat parameter layout (include/rau.h) and the stacked activation store of the
// answering-unit unroll.  The store is tensor-major ([nHop][B][dim] per tensor kind) so that the
// weight-gradient products of all hops can be issued as ONE contraction over nHop*B rows.
#pragma once
#include "rau_common.cuh"

constexpr int RAU_SP_ALIGN = 16;   // spatial pad granularity (UMMA N % 16, 16-byte bf16 rows)
static inline int rau_sp(int S) { return (S + RAU_SP_ALIGN - 1) / RAU_SP_ALIGN * RAU_SP_ALIGN; }

#define RAU_MULT_FIELDS(X) \
  X(Wq) X(bq) X(Wh) X(bh) X(Wi) X(bi) X(Wqa) X(bqa) X(Wa) X(ba) X(ws) X(Wm) X(bm) X(Wp) X(bp) \
  X(Wx) X(bx) X(Whh) X(bhh) X(Wo) X(bo) X(Ws) X(bso) X(wd) X(bs) X(bd)

template <typename P>
struct MultT {
#define X(n) P n;
  RAU_MULT_FIELDS(X)
#undef X
};

struct ParamEntry { const char* name; int64_t off; int64_t rows, cols; };

// fills entries for a group, returns total floats
int64_t rau_layout(const rau_config* cfg, int group, std::vector<ParamEntry>* out);

template <typename P, typename F>
static inline MultT<P> mult_views(const rau_config* cfg, F* flat) {
  std::vector<ParamEntry> e;
  rau_layout(cfg, 2, &e);
  MultT<P> m;
  int i = 0;
#define X(n) m.n = flat + e[i++].off;
  RAU_MULT_FIELDS(X)
#undef X
  return m;
}

struct RnnLayerOff { int64_t Wi, bi, Wh, bh; int in; };
static inline void rnn_offsets(const rau_config* cfg, RnnLayerOff out[4]) {
  int64_t off = 0;
  for (int L = 0; L < cfg->nlayer && L < 4; ++L) {
    const int in = L == 0 ? cfg->embed : cfg->Hq;
    out[L].in = in;
    out[L].Wi = off; off += (int64_t)4 * cfg->Hq * in;
    out[L].bi = off; off += 4 * cfg->Hq;
    out[L].Wh = off; off += (int64_t)4 * cfg->Hq * cfg->Hq;
    out[L].bh = off; off += 4 * cfg->Hq;
  }
}
