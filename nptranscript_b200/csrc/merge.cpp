// merge.cpp — host-side merge of per-rank cluster / sub-cluster tables by key (multi-GPU end of chromosome).
// Input: the table part of every rank's npt_results (first-read indices already global).  Output: one merged table set
// with global first-appearance ids (CigarClusters.java:83, CigarCluster.java:660) plus, for every rank, the maps from its
// local cluster ids / sub-cluster positions to the global ones.  Pure host code (hash maps over int vectors); the device
// side of the merge — permuting coverage rows into global order and summing them — is npt_remap_clusters + the caller's
// all-reduce.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/npt.h"

namespace {
struct VecHash {
  size_t operator()(const std::vector<int32_t>& v) const {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ v.size();
    for (int32_t x : v) { h ^= (uint32_t)x; h *= 0x100000001B3ull; h ^= h >> 29; }
    return (size_t)h;
  }
};
struct Merged {
  std::vector<int64_t> cl_first, sub_first, sub_sums;
  std::vector<uint32_t> cl_key_off, cl_sub_off, sub_key_off, sub_sum_off;
  std::vector<int32_t> cl_key_tok, cl_read_count, sub_key, sub_count, sub_break_sum, small, pair[5];
  std::vector<uint8_t> sub_flags;
  std::vector<std::vector<int32_t>> cmap, smap;
  npt_results res{};
};
inline uint64_t key_hash(int cl, const int32_t* k, size_t n) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uint32_t)cl << 32) ^ n;
  for (size_t i = 0; i < n; i++) { h ^= (uint32_t)k[i]; h *= 0x100000001B3ull; h ^= h >> 29; }
  return h;
}
}  // namespace

extern "C" {

// tables[r]: table fields of rank r's npt_results (per-read fields are ignored).  Returns an opaque handle.
void* npt_merge_tables(int nranks, const npt_results* const* tables, int num_sources) {
  const int S = num_sources;
  // ---- clusters (few): map of token vectors
  std::unordered_map<std::vector<int32_t>, int, VecHash> cl_index;
  std::vector<std::vector<int32_t>> cl_keys;
  std::vector<int64_t> cl_first;
  std::vector<std::vector<int>> cl_where(nranks);
  size_t total_subs = 0;
  for (int r = 0; r < nranks; r++) {
    const npt_results& t = *tables[r];
    total_subs += (size_t)t.n_subs;
    cl_where[r].resize(t.n_clusters);
    for (int c = 0; c < t.n_clusters; c++) {
      std::vector<int32_t> key(t.cl_key_tok + t.cl_key_off[c], t.cl_key_tok + t.cl_key_off[c + 1]);
      auto it = cl_index.find(key);
      int ci;
      if (it == cl_index.end()) { ci = (int)cl_keys.size(); cl_index.emplace(key, ci); cl_keys.push_back(std::move(key)); cl_first.push_back(t.cl_first_read[c]); }
      else { ci = it->second; cl_first[ci] = std::min(cl_first[ci], t.cl_first_read[c]); }
      cl_where[r][c] = ci;
    }
  }
  // ---- sub-clusters (many): flat accumulators + open addressing on a 64-bit hash of (cluster, key), keys verified
  struct Acc { int cl; int rank; uint32_t g; int64_t first; uint8_t flags; int32_t break_sum; uint32_t sum_off, nsum; };
  std::vector<Acc> acc; acc.reserve(total_subs);
  std::vector<int64_t> acc_count; acc_count.reserve(total_subs * S);
  std::vector<int64_t> acc_sums; acc_sums.reserve(total_subs * 6);
  size_t cap = 16; while (cap < total_subs * 2 + 16) cap <<= 1;
  std::vector<int> slots(cap, -1);
  std::vector<std::vector<int>> where(nranks);
  for (int r = 0; r < nranks; r++) {
    const npt_results& t = *tables[r];
    where[r].resize(t.n_subs);
    for (int c = 0; c < t.n_clusters; c++) {
      const int ci = cl_where[r][c];
      for (uint32_t g = t.cl_sub_off[c]; g < t.cl_sub_off[c + 1]; g++) {
        const int32_t* k = t.sub_key + t.sub_key_off[g];
        const size_t nk = t.sub_key_off[g + 1] - t.sub_key_off[g];
        const int64_t* sums = t.sub_sums + t.sub_sum_off[g];
        const uint32_t nsum = t.sub_sum_off[g + 1] - t.sub_sum_off[g];
        size_t pos = (size_t)key_hash(ci, k, nk) & (cap - 1);
        int a = -1;
        for (;; pos = (pos + 1) & (cap - 1)) {
          const int s = slots[pos];
          if (s < 0) break;
          const Acc& A = acc[s];
          const npt_results& tr = *tables[A.rank];
          const size_t nk2 = tr.sub_key_off[A.g + 1] - tr.sub_key_off[A.g];
          if (A.cl == ci && nk2 == nk && memcmp(tr.sub_key + tr.sub_key_off[A.g], k, nk * 4) == 0) { a = s; break; }
        }
        if (a < 0) {
          a = (int)acc.size();
          slots[pos] = a;
          acc.push_back({ci, r, g, t.sub_first_read[g], t.sub_flags[g], t.sub_break_sum[g], (uint32_t)acc_sums.size(), nsum});
          for (int j = 0; j < S; j++) acc_count.push_back(t.sub_count[(size_t)g * S + j]);
          acc_sums.insert(acc_sums.end(), sums, sums + nsum);
        } else {
          Acc& A = acc[a];
          if (t.sub_first_read[g] < A.first) { A.first = t.sub_first_read[g]; A.flags = t.sub_flags[g]; }
          A.break_sum += t.sub_break_sum[g];
          for (int j = 0; j < S; j++) acc_count[(size_t)a * S + j] += t.sub_count[(size_t)g * S + j];
          if (A.nsum == nsum) for (uint32_t j = 0; j < nsum; j++) acc_sums[A.sum_off + j] += sums[j];
        }
        where[r][g] = a;
      }
    }
  }
  // ---- global ids: clusters by first index; subs by (cluster id, first index)
  const int nC = (int)cl_keys.size();
  std::vector<int> order(nC), gid(nC);
  std::iota(order.begin(), order.end(), 0);
  std::sort(order.begin(), order.end(), [&](int a, int b) { return cl_first[a] < cl_first[b]; });
  for (int i = 0; i < nC; i++) gid[order[i]] = i;
  const int nS = (int)acc.size();
  std::vector<int> sorder(nS);
  std::iota(sorder.begin(), sorder.end(), 0);
  std::sort(sorder.begin(), sorder.end(), [&](int a, int b) {
    const int ga = gid[acc[a].cl], gb = gid[acc[b].cl];
    return ga != gb ? ga < gb : acc[a].first < acc[b].first;
  });
  Merged* M = new Merged();
  M->cl_key_off.push_back(0); M->sub_key_off.push_back(0); M->sub_sum_off.push_back(0);
  for (int i = 0; i < nC; i++) {
    M->cl_first.push_back(cl_first[order[i]]);
    M->cl_key_tok.insert(M->cl_key_tok.end(), cl_keys[order[i]].begin(), cl_keys[order[i]].end());
    M->cl_key_off.push_back((uint32_t)M->cl_key_tok.size());
  }
  M->cl_sub_off.assign(nC + 1, 0);
  M->cl_read_count.assign((size_t)nC * S, 0);
  std::vector<int> k_of_acc(nS);
  for (int p = 0; p < nS; p++) {
    const int a = sorder[p];
    const Acc& A = acc[a];
    const int g = gid[A.cl];
    M->cl_sub_off[g + 1]++;
    const npt_results& tr = *tables[A.rank];
    M->sub_first.push_back(A.first); M->sub_flags.push_back(A.flags); M->sub_break_sum.push_back(A.break_sum);
    M->sub_key.insert(M->sub_key.end(), tr.sub_key + tr.sub_key_off[A.g], tr.sub_key + tr.sub_key_off[A.g + 1]);
    M->sub_key_off.push_back((uint32_t)M->sub_key.size());
    M->sub_sums.insert(M->sub_sums.end(), acc_sums.begin() + A.sum_off, acc_sums.begin() + A.sum_off + A.nsum);
    M->sub_sum_off.push_back((uint32_t)M->sub_sums.size());
    for (int j = 0; j < S; j++) { const int64_t v = acc_count[(size_t)a * S + j]; M->sub_count.push_back((int32_t)v); M->cl_read_count[(size_t)g * S + j] += (int32_t)v; }
  }
  for (int g = 0; g < nC; g++) M->cl_sub_off[g + 1] += M->cl_sub_off[g];
  for (int p = 0; p < nS; p++) k_of_acc[sorder[p]] = p - (int)M->cl_sub_off[gid[acc[sorder[p]].cl]];
  M->cmap.resize(nranks); M->smap.resize(nranks);
  for (int r = 0; r < nranks; r++) {
    for (int ci : cl_where[r]) M->cmap[r].push_back(gid[ci]);
    for (int a : where[r]) M->smap[r].push_back(k_of_acc[a]);
  }
  // break-point pairs and small counters add up
  struct PK { int32_t a, b, c, d; bool operator<(const PK& o) const { return std::tie(a, b, c, d) < std::tie(o.a, o.b, o.c, o.d); } };
  std::vector<std::pair<PK, int64_t>> pairs;
  int n_orf = tables[0]->n_orf;
  M->small.assign((size_t)S * (1 + 2 * n_orf), 0);
  for (int r = 0; r < nranks; r++) {
    const npt_results& t = *tables[r];
    for (int64_t i = 0; i < t.n_pairs; i++) pairs.push_back({{t.pair_src[i], t.pair_level[i], t.pair_start[i], t.pair_end[i]}, t.pair_count[i]});
    for (int j = 0; j < S; j++) M->small[j] += t.total_counts[j];
    for (int j = 0; j < S * n_orf; j++) { M->small[S + j] += t.spliced[j]; M->small[S + S * n_orf + j] += t.unspliced[j]; }
  }
  std::sort(pairs.begin(), pairs.end(), [](const auto& x, const auto& y) { return x.first < y.first; });
  for (size_t i = 0; i < pairs.size(); i++) {
    const PK& k = pairs[i].first;
    if (!M->pair[0].empty() && M->pair[0].back() == k.a && M->pair[1].back() == k.b && M->pair[2].back() == k.c && M->pair[3].back() == k.d) { M->pair[4].back() += (int32_t)pairs[i].second; continue; }
    M->pair[0].push_back(k.a); M->pair[1].push_back(k.b); M->pair[2].push_back(k.c); M->pair[3].push_back(k.d); M->pair[4].push_back((int32_t)pairs[i].second);
  }
  npt_results& R = M->res;
  R.n_clusters = nC; R.n_subs = (int32_t)M->sub_first.size(); R.n_orf = n_orf;
  R.cl_first_read = M->cl_first.data(); R.cl_key_off = M->cl_key_off.data(); R.cl_key_tok = M->cl_key_tok.data(); R.cl_read_count = M->cl_read_count.data();
  R.cl_sub_off = M->cl_sub_off.data(); R.sub_first_read = M->sub_first.data(); R.sub_key_off = M->sub_key_off.data(); R.sub_key = M->sub_key.data();
  R.sub_count = M->sub_count.data(); R.sub_break_sum = M->sub_break_sum.data(); R.sub_sum_off = M->sub_sum_off.data(); R.sub_sums = M->sub_sums.data();
  R.sub_flags = M->sub_flags.data(); R.total_counts = M->small.data(); R.spliced = M->small.data() + S; R.unspliced = M->small.data() + S + S * n_orf;
  R.n_pairs = (int64_t)M->pair[0].size();
  R.pair_src = M->pair[0].data(); R.pair_level = M->pair[1].data(); R.pair_start = M->pair[2].data(); R.pair_end = M->pair[3].data(); R.pair_count = M->pair[4].data();
  return M;
}

const npt_results* npt_merged_results(void* h) { return &((Merged*)h)->res; }
const int32_t* npt_merged_cluster_map(void* h, int rank) { return ((Merged*)h)->cmap[rank].data(); }
const int32_t* npt_merged_sub_map(void* h, int rank) { return ((Merged*)h)->smap[rank].data(); }
void npt_merge_free(void* h) { delete (Merged*)h; }

}  // extern "C"
