// Seed chaining and chain filtering, one read per task.
//
// Restates bwa/bwamem.c:190-211 (test_and_merge), :213-233 (mem_chain_weight), :251-315 (mem_chain)
// and :327-385 (mem_chain_flt).  The reference keeps chains in a B-tree keyed by the chain's first
// reference position only (bwamem.c:186-187; kbtree.h with t = 5 for a 40-byte key at node size
// 512).  Duplicate keys do occur, and which duplicate a lookup returns as well as the final in-order
// listing depend on the node splits, so the same tree discipline is rebuilt here:
// lower-bound probe inside a node (kbtree.h:118-131), predecessor search (:152-168), insertion
// after the probe position with pre-emptive splitting of full children (:174-224) and in-order
// traversal (:336-358) -- over integer chain ids in a flat node pool instead of pointers.
#pragma once
#include "b2_types.h"
#include "b2_sort.cuh"

#define B2_BT_T 5
#define B2_BT_MAXK (2 * B2_BT_T - 1)

struct B2BtNode {
  int32_t n, is_internal;
  int32_t key[B2_BT_MAXK];
  int32_t ptr[B2_BT_MAXK + 1];
};

struct B2Btree {
  B2BtNode* nodes;
  const B2Chain* ch;
  int n_nodes, cap, root, n_keys, overflow;
};

B2_HD inline int b2_bt_new(B2Btree& b, int internal) {
  if (b.n_nodes >= b.cap) { b.overflow = 1; return b.cap - 1; }
  B2BtNode& z = b.nodes[b.n_nodes];
  z.n = 0; z.is_internal = internal;
  for (int i = 0; i <= B2_BT_MAXK; ++i) z.ptr[i] = -1;
  return b.n_nodes++;
}

B2_HD inline void b2_bt_init(B2Btree& b, B2BtNode* pool, int cap, const B2Chain* ch) {
  b.nodes = pool; b.cap = cap; b.ch = ch; b.n_nodes = 0; b.n_keys = 0; b.overflow = 0;
  b.root = b2_bt_new(b, 0);
}

// position of the last key <= pos (first one among equals); *r == 0 iff an equal key was hit
B2_HD inline int b2_bt_probe(const B2Btree& b, const B2BtNode& x, int64_t pos, int* r) {
  if (x.n == 0) return -1;
  int begin = 0, end = x.n;
  while (begin < end) {
    int mid = (begin + end) >> 1;
    if (b.ch[x.key[mid]].pos < pos) begin = mid + 1; else end = mid;
  }
  if (begin == x.n) { *r = 1; return x.n - 1; }
  *r = (b.ch[x.key[begin]].pos < pos) - (pos < b.ch[x.key[begin]].pos);
  if (*r < 0) --begin;
  return begin;
}

// chain id of the closest chain at or before pos, or -1
B2_HD inline int b2_bt_lower(const B2Btree& b, int64_t pos) {
  int lower = -1, x = b.root, r = 0;
  while (x >= 0) {
    const B2BtNode& nd = b.nodes[x];
    int i = b2_bt_probe(b, nd, pos, &r);
    if (i >= 0 && r == 0) return nd.key[i];
    if (i >= 0) lower = nd.key[i];
    if (!nd.is_internal) return lower;
    x = nd.ptr[i + 1];
  }
  return lower;
}

B2_HD inline void b2_bt_split(B2Btree& b, int xi, int i, int yi) {
  int zi = b2_bt_new(b, b.nodes[yi].is_internal);
  B2BtNode &x = b.nodes[xi], &y = b.nodes[yi], &z = b.nodes[zi];
  z.n = B2_BT_T - 1;
  for (int j = 0; j < B2_BT_T - 1; ++j) z.key[j] = y.key[j + B2_BT_T];
  if (y.is_internal) for (int j = 0; j < B2_BT_T; ++j) z.ptr[j] = y.ptr[j + B2_BT_T];
  y.n = B2_BT_T - 1;
  for (int j = x.n; j > i; --j) x.ptr[j + 1] = x.ptr[j];
  x.ptr[i + 1] = zi;
  for (int j = x.n - 1; j >= i; --j) x.key[j + 1] = x.key[j];
  x.key[i] = y.key[B2_BT_T - 1];
  ++x.n;
}

B2_HD inline void b2_bt_put(B2Btree& b, int cid) {
  const int64_t pos = b.ch[cid].pos;
  ++b.n_keys;
  int r = b.root, dummy;
  if (b.nodes[r].n == B2_BT_MAXK) {
    int s = b2_bt_new(b, 1);
    b.root = s;
    b.nodes[s].ptr[0] = r;
    b2_bt_split(b, s, 0, r);
    r = s;
  }
  int x = r;
  for (;;) {
    B2BtNode& nd = b.nodes[x];
    if (!nd.is_internal) {
      int i = b2_bt_probe(b, nd, pos, &dummy);
      for (int j = nd.n - 1; j > i; --j) nd.key[j + 1] = nd.key[j];
      nd.key[i + 1] = cid;
      ++nd.n;
      return;
    }
    int i = b2_bt_probe(b, nd, pos, &dummy) + 1;
    if (b.nodes[nd.ptr[i]].n == B2_BT_MAXK) {
      b2_bt_split(b, x, i, nd.ptr[i]);
      const B2BtNode& nd2 = b.nodes[x];
      int64_t kp = b.ch[nd2.key[i]].pos;
      if (((kp < pos) - (pos < kp)) > 0) ++i;
    }
    x = b.nodes[x].ptr[i];
    if (b.overflow) return;
  }
}

// in-order listing of chain ids into out[]; returns the count
B2_HD inline int b2_bt_traverse(const B2Btree& b, int* out) {
  int sx[24], si[24], sp = 0, n = 0;
  sx[0] = b.root; si[0] = 0;
  for (;;) {
    while (sx[sp] >= 0 && si[sp] <= b.nodes[sx[sp]].n) {
      const B2BtNode& nd = b.nodes[sx[sp]];
      int child = nd.is_internal ? nd.ptr[si[sp]] : -1;
      ++sp;
      sx[sp] = child; si[sp] = 0;
    }
    --sp;
    if (sp < 0) break;
    if (sx[sp] >= 0 && si[sp] < b.nodes[sx[sp]].n) out[n++] = b.nodes[sx[sp]].key[si[sp]];
    ++si[sp];
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
B2_HD inline int b2_test_and_merge(const B2Opt& opt, int64_t l_pac, B2Chain& c, B2Seed* seeds, int pi, int seed_rid) {
  const B2Seed& p = seeds[pi];
  const B2Seed& first = seeds[c.seed_head];
  const B2Seed& last = seeds[c.seed_tail];
  int64_t qend = last.qbeg + last.len, rend = last.rbeg + last.len;
  if (seed_rid != c.rid) return 0;
  if (p.qbeg >= first.qbeg && p.qbeg + p.len <= qend && p.rbeg >= first.rbeg && p.rbeg + p.len <= rend) return 1;
  if ((last.rbeg < l_pac || first.rbeg < l_pac) && p.rbeg >= l_pac) return 0;
  int64_t x = p.qbeg - last.qbeg, y = p.rbeg - last.rbeg;
  if (y >= 0 && x - y <= opt.w && y - x <= opt.w && x - last.len < opt.max_chain_gap && y - last.len < opt.max_chain_gap) {
    seeds[c.seed_tail].next = pi;
    c.seed_tail = pi;
    seeds[pi].next = -1;
    ++c.n;
    return 1;
  }
  return 0;
}

// weight over a contiguous seed array
B2_HD inline int b2_chain_weight(const B2Seed* s, int n) {
  int64_t end = 0;
  int w = 0, tmp;
  for (int j = 0; j < n; ++j) {
    if (s[j].qbeg >= end) w += s[j].len;
    else if (s[j].qbeg + s[j].len > end) w += (int)(s[j].qbeg + s[j].len - end);
    end = end > s[j].qbeg + s[j].len ? end : s[j].qbeg + s[j].len;
  }
  tmp = w; w = 0; end = 0;
  for (int j = 0; j < n; ++j) {
    if (s[j].rbeg >= end) w += s[j].len;
    else if (s[j].rbeg + s[j].len > end) w += (int)(s[j].rbeg + s[j].len - end);
    end = end > s[j].rbeg + s[j].len ? end : s[j].rbeg + s[j].len;
  }
  w = w < tmp ? w : tmp;
  return w < 1 << 30 ? w : (1 << 30) - 1;
}

struct B2ChainWLt { B2_HD bool operator()(const B2Chain& a, const B2Chain& b) const { return a.w > b.w; } };

// Build chains for one read.
//   intv[n_intv]   sorted seeding intervals
//   raw[n_raw]     one entry per sampled suffix-array position in task order: rbeg/qbeg/len/score set,
//                  `next` holds the contig id from bns_intv2rid (negative: seed is skipped)
//   tree_ch[n_raw] chain pool while chaining; ord[n_raw] chains in key order -> filtered prefix
//   seeds_out[n_raw] seeds regrouped per surviving chain (chain.seed_head = offset)
//   order[n_raw]   int scratch
// Returns the number of chains kept after filtering; *overflow set if the node pool was exhausted.
B2_HD inline int b2_chain_read(const B2Opt& opt, int64_t l_pac, const B2Ann* anns, int len, const B2Intv* intv, int n_intv,
                               B2Seed* raw, int n_raw, B2Chain* tree_ch, B2Chain* ord, B2Seed* seeds_out, int* order,
                               B2BtNode* nodes, int node_cap, int* overflow) {
  if (len < opt.min_seed_len || n_raw == 0) return 0;
  B2Btree tree;
  b2_bt_init(tree, nodes, node_cap, tree_ch);
  int b = 0, e = 0, l_rep = 0;
  for (int i = 0; i < n_intv; ++i) {
    int sb = (int)(intv[i].info >> 32), se = (int)(uint32_t)intv[i].info;
    if (intv[i].x[2] <= (uint64_t)opt.max_occ) continue;
    if (sb > e) { l_rep += e - b; b = sb; e = se; }
    else e = e > se ? e : se;
  }
  l_rep += e - b;
  int n_ch = 0;
  for (int si = 0; si < n_raw; ++si) {
    int rid = raw[si].next;
    raw[si].next = -1;
    if (rid < 0) continue;
    int to_add = 0;
    if (tree.n_keys) {
      int lower = b2_bt_lower(tree, raw[si].rbeg);
      if (lower < 0 || !b2_test_and_merge(opt, l_pac, tree_ch[lower], raw, si, rid)) to_add = 1;
    } else to_add = 1;
    if (to_add) {
      B2Chain& c = tree_ch[n_ch];
      c.n = 1; c.first = -1; c.rid = rid; c.w = 0; c.kept = 0;
      c.is_alt = anns[rid].is_alt ? 1 : 0;
      c.frac_rep = 0.f; c.pos = raw[si].rbeg;
      c.seed_head = c.seed_tail = si;
      b2_bt_put(tree, n_ch);
      ++n_ch;
      if (tree.overflow) { *overflow = 1; return 0; }
    }
  }
  int n = b2_bt_traverse(tree, order);
  // regroup: chains in key order, seeds contiguous per chain
  int so = 0;
  const float frac_rep = (float)l_rep / len;
  for (int i = 0; i < n; ++i) {
    B2Chain c = tree_ch[order[i]];
    int off = so;
    for (int s = c.seed_head; s >= 0; s = raw[s].next) seeds_out[so++] = raw[s];
    c.seed_head = off; c.seed_tail = so - 1;
    c.frac_rep = frac_rep;
    ord[i] = c;
  }
  // ---- filter (mem_chain_flt) ----
  if (n == 0) return 0;
  int k = 0;
  for (int i = 0; i < n; ++i) {
    B2Chain& c = ord[i];
    c.first = -1; c.kept = 0;
    c.w = (uint32_t)b2_chain_weight(seeds_out + c.seed_head, c.n);
    if ((int)c.w < opt.min_chain_weight) continue;
    ord[k++] = c;
  }
  n = k;
  if (n == 0) return 0;  // (the reference would touch a[0] here; unreachable with min_chain_weight <= every weight)
  b2_introsort(ord, (long)n, B2ChainWLt());
#define B2_CB(c) (seeds_out[(c).seed_head].qbeg)
#define B2_CE(c) (seeds_out[(c).seed_tail].qbeg + seeds_out[(c).seed_tail].len)
  int* kept_idx = order;  // reuse
  int n_kept = 0;
  ord[0].kept = 3;
  kept_idx[n_kept++] = 0;
  for (int i = 1; i < n; ++i) {
    int large_ovlp = 0, kk;
    for (kk = 0; kk < n_kept; ++kk) {
      int j = kept_idx[kk];
      int b_max = B2_CB(ord[j]) > B2_CB(ord[i]) ? B2_CB(ord[j]) : B2_CB(ord[i]);
      int e_min = B2_CE(ord[j]) < B2_CE(ord[i]) ? B2_CE(ord[j]) : B2_CE(ord[i]);
      if (e_min > b_max && (!ord[j].is_alt || ord[i].is_alt)) {
        int li = B2_CE(ord[i]) - B2_CB(ord[i]);
        int lj = B2_CE(ord[j]) - B2_CB(ord[j]);
        int min_l = li < lj ? li : lj;
        if (e_min - b_max >= min_l * opt.mask_level && min_l < opt.max_chain_gap) {
          large_ovlp = 1;
          if (ord[j].first < 0) ord[j].first = i;
          if (ord[i].w < ord[j].w * opt.drop_ratio && (int)(ord[j].w - ord[i].w) >= opt.min_seed_len << 1) break;
        }
      }
    }
    if (kk == n_kept) {
      kept_idx[n_kept++] = i;
      ord[i].kept = large_ovlp ? 2 : 3;
    }
  }
#undef B2_CB
#undef B2_CE
  for (int i = 0; i < n_kept; ++i) {
    const B2Chain& c = ord[kept_idx[i]];
    if (c.first >= 0) ord[c.first].kept = 1;
  }
  int i;
  for (i = k = 0; i < n; ++i) {
    if (ord[i].kept == 0 || ord[i].kept == 3) continue;
    if (++k >= opt.max_chain_extend) break;
  }
  for (; i < n; ++i)
    if (ord[i].kept < 3) ord[i].kept = 0;
  for (i = k = 0; i < n; ++i)
    if (ord[i].kept != 0) ord[k++] = ord[i];
  return k;
}
