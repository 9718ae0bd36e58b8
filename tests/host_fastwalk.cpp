// host_fastwalk.cpp — TEST INFRASTRUCTURE.  Compiles the lane logic of nptranscript_b200/csrc/fastwalk.cuh (the same
// header the CUDA kernel includes) with g++ and runs it one read at a time, staging the CIGAR / base tiles exactly like
// fast_walk_kernel does.  tests/test_fastwalk_host.py compares what it writes with the oracle, so the task / ring / entry
// logic is checked on the CPU box; the GPU tests then check the kernel around it.  Never loaded by the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../nptranscript_b200/csrc/fastwalk.cuh"

using namespace npt;

template <bool RECORD>
static void run_all(const DevCfg& c, const BatchDev& b) {
  std::vector<uint32_t> tile(FW_TILE_STRIDE + 4), ring(FW_RING + FW_RING_PAD + FW_BRK_BUF, 0);
  uint32_t* cigT = tile.data();
  uint32_t* seqT = tile.data() + FW_CIG_TILE;
  for (long long r = 0; r < b.n; r++) {
    FwLane L;
    memset(&L, 0, sizeof L);
    fw_start_read(L, b, r);
    while (L.flags & FWF_ACTIVE) {
      const bool live = (L.flags & (FWF_ACTIVE | FWF_SLOW)) == FWF_ACTIVE;
      const unsigned ncs = L.ck & ~3u;
      if (live && ncs != L.cs && L.ck < L.cend) {
        L.cs = ncs;
        for (int i = 0; i < FW_CIG_TILE; i++) cigT[i] = ((long long)ncs + i < b.cigar_words) ? b.cigar[ncs + i] : 0u;
      }
      const long long ntb = (long long)((L.so + (unsigned long long)(L.readPos >> 1)) & ~15ull);
      if (live && ntb != L.tb) {
        L.tb = ntb; L.tnib0 = (int)(2 * (ntb - (long long)L.so));
        unsigned char* dst = (unsigned char*)seqT;
        for (int i = 0; i < FW_SEQ_TILE * 4; i++) dst[i] = (ntb + i < b.seq_bytes) ? b.seq4[ntb + i] : 0;
      }
      const int ringLimit = (L.flushE << 5) + FW_RING * 8 - 2 * FW_PIECE - 2;
      for (int it = 0; it < FW_R; it++) fw_step<RECORD>(L, c, b, cigT, seqT, c.refg_words, ring.data(), 1, live, ringLimit);
      fw_round_end<RECORD>(L, c, b, cigT, ring.data(), 1);
    }
  }
}

extern "C" int hfw_run(int T, int G, int record, const uint32_t* refg, int refg_nwords, long long n, const int* pos, const uint32_t* cigar_off,
                       const uint32_t* cigar, const unsigned long long* seq_off, const uint8_t* seq4, int* nbreaks, int* break_arena,
                       int* read_st, int* read_en, int* maps_sum, int* err_sum, unsigned* rflags, unsigned* break_loc, uint4* fhdrA,
                       unsigned* fmeta, unsigned* frare, uint4* fent, unsigned* slow_list, unsigned* bctr, unsigned break_cap) {
  DevCfg c;
  memset(&c, 0, sizeof c);
  c.T = T; c.G = G; c.stride = G + 2; c.record_depth = record; c.refg_words = refg; c.refg_nwords = refg_nwords;
  BatchDev b;
  memset(&b, 0, sizeof b);
  b.n = n; b.pos = pos; b.cigar_off = cigar_off; b.cigar = cigar; b.seq_off = seq_off; b.seq4 = seq4;
  b.cigar_words = cigar_off[n]; b.seq_bytes = (long long)seq_off[n];
  b.nbreaks = nbreaks; b.break_arena = break_arena; b.read_st = read_st; b.read_en = read_en; b.maps_sum = maps_sum; b.err_sum = err_sum;
  b.rflags = rflags; b.break_loc = break_loc; b.fhdrA = fhdrA; b.fmeta = fmeta; b.frare = frare; b.fent = fent; b.slow_list = slow_list; b.bctr = bctr; b.fast = 1; b.break_cap = break_cap;
  if (record) run_all<true>(c, b); else run_all<false>(c, b);
  return (int)bctr[7];
}
