// comm.inl — the multi-GPU half of libnpt.so (included at the end of npt.cu): collectives inside the library, and the
// in-process multi-device context (npt_mg_*) a single-threaded host such as the Java shim drives.
//
// What the reference does here: one IdentityProfileHolder per chromosome receives every record of every input BAM
// (J/cluster/IdentityProfileHolder.java:170-196) and owns ONE CigarClusters table (J/cluster/CigarClusters.java:75-102).
// Sharded over GPUs the reads split into contiguous ranges of the submit order; nothing is exchanged while reads are walked.
// At end of chromosome there is one real exchange step:
//   1. every rank packs the KEYS of its tables (npt_export_keys), the 16-int headers are all-gathered so that every rank
//      knows every buffer's size, and the buffers travel with one group of ncclBroadcast calls (an all-gather with
//      different sizes per rank);
//   2. every rank inserts the other ranks' keys (import kernels enqueued back to back on its compute stream, no host
//      round trip in between);
//   3. npt_end_chrom numbers clusters / sub-clusters / break-point pairs — now identically on every rank;
//   4. the counters, in that common order, are summed in place with ncclAllReduce (coverage rows, ORF counters,
//      sub-cluster counts, break sums, pair counts).
// NCCL is loaded with dlopen("libnccl.so.2") the first time a communicator is needed, so that a single-GPU host needs no
// NCCL at all and a process that already carries torch's bundled NCCL shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <mutex>
#include <thread>

namespace {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string why;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = nullptr;
    const char* env = getenv("NPT_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) if (nm && !h) h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
    bool all = true;
    auto sym = [&](auto& fp, const char* nm) { fp = reinterpret_cast<std::remove_reference_t<decltype(fp)>>(dlsym(h, nm)); if (!fp) { all = false; api.why = std::string("missing symbol ") + nm; } };
    sym(api.GetUniqueId, "ncclGetUniqueId"); sym(api.CommInitRank, "ncclCommInitRank"); sym(api.CommInitAll, "ncclCommInitAll");
    sym(api.CommDestroy, "ncclCommDestroy"); sym(api.GroupStart, "ncclGroupStart"); sym(api.GroupEnd, "ncclGroupEnd");
    sym(api.Broadcast, "ncclBroadcast"); sym(api.AllReduce, "ncclAllReduce"); sym(api.AllGather, "ncclAllGather");
    sym(api.GetErrorString, "ncclGetErrorString");
    api.ok = all;
  });
  return api;
}

// several contexts of ONE process on one device (tests on a single-GPU box): the same protocol with device-to-device copies
// and an add kernel in place of NCCL.  Shared by the contexts of an npt_mg.
struct LocalGroup {
  std::vector<npt_ctx*> members;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long generation = 0;
  bool aborted = false;   // a member failed: nobody waits for it any more (until the next chromosome)
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    const long long g = generation;
    if (++arrived == (int)members.size()) { arrived = 0; generation++; cv.notify_all(); }
    else cv.wait(lk, [&] { return generation != g || aborted; });
    return !aborted;
  }
  void abort() { std::lock_guard<std::mutex> lk(mu); aborted = true; cv.notify_all(); }
  void reset() { std::lock_guard<std::mutex> lk(mu); aborted = false; arrived = 0; }
};

__global__ void local_sum_kernel(int* const* ptrs, int n, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    int s = 0;
    for (int k = 0; k < n; k++) s += ptrs[k][i];
    for (int k = 0; k < n; k++) ptrs[k][i] = s;
  }
}

}  // namespace

struct NptComm {
  ncclComm_t nc = nullptr;       // NCCL communicator of this rank (null in a LocalGroup)
  LocalGroup* local = nullptr;   // not owned
};

#define NCK(call)                                                                                                     \
  do {                                                                                                                \
    ncclResult_t r__ = (call);                                                                                        \
    if (r__ != ncclSuccess) return fail(ctx, NPT_ECUDA, "%s: %s (%s:%d)", #call, nccl().GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

static void comm_release(npt_ctx* ctx) {
  if (ctx->comm) {
    if (ctx->comm->nc && nccl().ok) nccl().CommDestroy(ctx->comm->nc);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
  ctx->d_kx_all.release(); ctx->d_kx_hdrs.release();
  ctx->comm_rank = 0; ctx->comm_world = 1;
}

// ---- the collective end of chromosome of one rank
static int end_chrom_all(npt_ctx* ctx) {
  if (!ctx) return NPT_EINVAL;
  if (!ctx->comm || ctx->comm_world <= 1) return npt_end_chrom(ctx);
  if (!ctx->in_chrom || ctx->finalized) return fail(ctx, NPT_ESTATE, "npt_end_chrom_all belongs after the last npt_submit of a chromosome");
  CK(cudaSetDevice(ctx->device));
  const int W = ctx->comm_world, me = ctx->comm_rank;
  NptComm& cm = *ctx->comm;
  cudaStream_t st = ctx->s_comp;
  cudaEvent_t e0, e1, e2, e3;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
  auto drop = [&] { cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2); cudaEventDestroy(e3); };
  // an error of one member of a local group releases the others from their barriers (they fail with NPT_ESTATE)
  auto bail = [&](int code) { drop(); if (cm.local) cm.local->abort(); return code; };
  auto sync_group = [&]() { return cm.local->barrier() ? NPT_OK : fail(ctx, NPT_ESTATE, "another context of the group failed"); };
  // 1. this rank's keys
  const int32_t* mine = nullptr;
  int64_t my_words = 0;
  int rc = npt_export_keys(ctx, &mine, &my_words);
  if (rc) return bail(rc);
  CK(cudaEventRecord(e0, st));
  // headers of all ranks -> host
  std::vector<int> hdrs((size_t)W * KX_HDR);
  CK(ctx->d_kx_hdrs.ensure((size_t)W * KX_HDR));
  if (cm.nc) {
    NCK(nccl().AllGather(mine, ctx->d_kx_hdrs.p, KX_HDR, ncclInt32, cm.nc, st));
    CK(cudaMemcpyAsync(hdrs.data(), ctx->d_kx_hdrs.p, hdrs.size() * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    if ((rc = sync_group())) return bail(rc);   // every member has exported (npt_export_keys waits for its stream)
    for (int r = 0; r < W; r++) memcpy(hdrs.data() + (size_t)r * KX_HDR, cm.local->members[r]->kx_hdr, KX_HDR * sizeof(int));
  }
  std::vector<long long> off(W + 1, 0);
  for (int r = 0; r < W; r++) {
    const int* h = hdrs.data() + (size_t)r * KX_HDR;
    if (h[0] != KX_MAGIC || h[13] < KX_HDR) return bail(fail(ctx, NPT_EINVAL, "rank %d sent no key buffer", r));
    off[r + 1] = off[r] + ((h[13] + 3) & ~3);   // 16-byte aligned sections
  }
  CK(ctx->d_kx_all.ensure((size_t)off[W]));
  // 2. the buffers: one broadcast per rank, all in one group
  if (cm.nc) {
    NCK(nccl().GroupStart());
    for (int r = 0; r < W; r++)
      NCK(nccl().Broadcast(mine, ctx->d_kx_all.p + off[r], (size_t)hdrs[(size_t)r * KX_HDR + 13], ncclInt32, r, cm.nc, st));
    NCK(nccl().GroupEnd());
  } else {
    for (int r = 0; r < W; r++)
      if (r != me)
        CK(cudaMemcpyAsync(ctx->d_kx_all.p + off[r], cm.local->members[r]->d_kx.p, (size_t)hdrs[(size_t)r * KX_HDR + 13] * 4, cudaMemcpyDeviceToDevice, st));
  }
  // 3. insert the other ranks' keys
  for (int r = 0; r < W; r++)
    if (r != me && (rc = import_keys_enqueue(ctx, ctx->d_kx_all.p + off[r], hdrs[(size_t)r * KX_HDR + 13], hdrs.data() + (size_t)r * KX_HDR))) return bail(rc);
  CK(cudaEventRecord(e1, st));
  if (!cm.nc) { CK(cudaStreamSynchronize(st)); if ((rc = sync_group())) return bail(rc); }   // nobody rewrites its d_kx before everyone has copied it
  // 4. global numbering
  if ((rc = npt_end_chrom(ctx))) return bail(rc);
  // 5. sum the counters in place
  npt_device_view v;
  if ((rc = npt_device_view_get(ctx, &v))) return bail(rc);
  struct Arr { void* p; long long n; bool wide; } arrs[6] = {{v.coverage, v.coverage_elems, false}, {v.small_counts, v.small_elems, false},
      {v.sub_count, v.sub_count_elems, false}, {v.sub_break_sum, v.sub_break_sum_elems, false}, {v.sub_sums, v.sub_sums_elems, true},
      {v.pair_count, v.pair_count_elems, false}};
  CK(cudaEventRecord(e2, st));
  if (cm.nc) {
    NCK(nccl().GroupStart());
    for (auto& a : arrs)
      if (a.p && a.n) NCK(nccl().AllReduce(a.p, a.p, (size_t)a.n, a.wide ? ncclInt64 : ncclInt32, ncclSum, cm.nc, st));
    NCK(nccl().GroupEnd());
  } else {
    if ((rc = sync_group())) return bail(rc);   // every member has finished npt_end_chrom (it waits for its stream)
    if (me == 0) {
      DBuf<int*> d_ptrs;
      CK(d_ptrs.ensure((size_t)W));
      for (int k = 0; k < 6; k++) {
        std::vector<int*> hp(W);
        long long cnt = -1;
        for (int r = 0; r < W; r++) {
          npt_device_view vr;
          if ((rc = npt_device_view_get(cm.local->members[r], &vr))) return bail(rc);
          void* ps[6] = {vr.coverage, vr.small_counts, vr.sub_count, vr.sub_break_sum, vr.sub_sums, vr.pair_count};
          const long long ns[6] = {vr.coverage_elems, vr.small_elems, vr.sub_count_elems, vr.sub_break_sum_elems, 2 * vr.sub_sums_elems, vr.pair_count_elems};
          hp[r] = (int*)ps[k];
          if (cnt >= 0 && cnt != ns[k]) return bail(fail(ctx, NPT_ECUDA, "internal: counter arrays differ in size between ranks"));
          cnt = ns[k];
        }
        if (cnt <= 0 || !hp[0]) continue;
        CK(cudaMemcpyAsync(d_ptrs.p, hp.data(), (size_t)W * sizeof(int*), cudaMemcpyHostToDevice, st));
        if (k == 4) {   // int64 sums: added as pairs of 32-bit halves would drop carries; use the 64-bit view
          // (arrays are small: one thread per element through the generic kernel on a long long view)
          std::vector<long long> acc((size_t)cnt / 2, 0), tmp((size_t)cnt / 2);
          CK(cudaStreamSynchronize(st));
          for (int r = 0; r < W; r++) { CK(cudaMemcpy(tmp.data(), hp[r], (size_t)cnt * 4, cudaMemcpyDeviceToHost)); for (size_t i = 0; i < acc.size(); i++) acc[i] += tmp[i]; }
          for (int r = 0; r < W; r++) CK(cudaMemcpy(hp[r], acc.data(), (size_t)cnt * 4, cudaMemcpyHostToDevice));
        } else {
          local_sum_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(d_ptrs.p, W, cnt);
          CK(cudaGetLastError());
          CK(cudaStreamSynchronize(st));
        }
      }
      d_ptrs.release();
    }
    if ((rc = sync_group())) return bail(rc);
  }
  CK(cudaEventRecord(e3, st));
  CK(cudaEventRecord(ctx->ev_c1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&ctx->kx_ms, e0, e1));
  CK(cudaEventElapsedTime(&ctx->reduce_ms, e2, e3));
  CK(cudaEventElapsedTime(&ctx->step_ms, ctx->ev_c0, ctx->ev_c1));
  drop();
  return NPT_OK;
}

// ================================================================================================= in-process multi-device
struct npt_mg {
  std::vector<npt_ctx*> ctx;
  std::vector<int> devices;
  npt_config cfg{};
  LocalGroup local;
  bool use_local = false;
  std::string err;
  int64_t user_base = 0, next_index = 0;
  // reads of (batch, device) in submit order: how the per-read results interleave
  std::vector<std::vector<int64_t>> part;     // part[batch][device]
  // rebased offsets of a device's share of a batch (pinned; ring of three: a slot is reused only after the batch that used
  // it has been retired by a later npt_submit of that device)
  struct Scratch {
    uint32_t* coff = nullptr; uint64_t* soff = nullptr; size_t cap = 0;
    uint32_t* widx = nullptr; size_t wcap = 0;      // rebased wide-element / exception indices of a packed share
    uint64_t* eidx = nullptr; size_t ecap = 0;
  };
  std::vector<Scratch> scratch;               // [device * 3 + slot]
  int64_t n_batches = 0;
  // merged results
  npt_results res{};
  std::vector<int32_t> r_i32[8];
  std::vector<uint32_t> r_flags;
  std::vector<uint64_t> r_boff;
  std::vector<int32_t> r_breaks;
};

namespace {

int mg_fail(npt_mg* m, int code, const char* fmt, ...) {
  char buf[640];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (m) m->err = buf; else g_create_error = buf;
  return code;
}

// run f(d) for every device on its own host thread (SURVEY §8b "Threading": one host thread per GPU); first error wins
template <typename F>
int mg_parallel(npt_mg* m, F f) {
  const int n = (int)m->ctx.size();
  std::vector<int> rcs(n, 0);
  if (n == 1) rcs[0] = f(0);
  else {
    std::vector<std::thread> th;
    for (int d = 0; d < n; d++) th.emplace_back([&, d] { rcs[d] = f(d); });
    for (auto& t : th) t.join();
  }
  for (int d = 0; d < n; d++)
    if (rcs[d]) return mg_fail(m, rcs[d], "device %d: %s", m->devices[d], m->ctx[d]->err.c_str());
  return NPT_OK;
}

}  // namespace

extern "C" {

int npt_comm_unique_id(void* id128) {
  if (!id128) return NPT_EINVAL;
  if (!nccl().ok) return fail(nullptr, NPT_ENODEVICE, "NCCL not available: %s", nccl().why.c_str());
  ncclUniqueId id;
  if (nccl().GetUniqueId(&id) != ncclSuccess) return fail(nullptr, NPT_ECUDA, "ncclGetUniqueId failed");
  memcpy(id128, &id, NPT_COMM_ID_BYTES);
  return NPT_OK;
}

int npt_comm_init(npt_ctx* ctx, const void* id128, int rank, int world) {
  if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, NPT_EINVAL, "bad argument");
  comm_release(ctx);
  if (world == 1) return NPT_OK;
  if (!nccl().ok) return fail(ctx, NPT_ENODEVICE, "NCCL not available: %s", nccl().why.c_str());
  CK(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, id128, NPT_COMM_ID_BYTES);
  ncclComm_t c = nullptr;
  NCK(nccl().CommInitRank(&c, world, id, rank));
  ctx->comm = new NptComm();
  ctx->comm->nc = c;
  ctx->comm_rank = rank; ctx->comm_world = world;
  return NPT_OK;
}

int npt_end_chrom_all(npt_ctx* ctx) { return end_chrom_all(ctx); }

int npt_last_exchange_ms(npt_ctx* ctx, float* key_exchange_ms, float* reduce_ms) {
  if (!ctx) return NPT_EINVAL;
  if (key_exchange_ms) *key_exchange_ms = ctx->kx_ms;
  if (reduce_ms) *reduce_ms = ctx->reduce_ms;
  return NPT_OK;
}

const char* npt_mg_last_error(const npt_mg* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

void npt_mg_destroy(npt_mg* m) {
  if (!m) return;
  for (auto* c : m->ctx) npt_destroy(c);
  for (auto& s : m->scratch) { if (s.coff) cudaFreeHost(s.coff); if (s.soff) cudaFreeHost(s.soff); if (s.widx) cudaFreeHost(s.widx); if (s.eidx) cudaFreeHost(s.eidx); }
  delete m;
}

int npt_mg_create(const npt_config* cfg, const int32_t* devices, int32_t n_devices, npt_mg** out) {
  if (!cfg || !devices || !out || n_devices < 1 || n_devices > 64) return mg_fail(nullptr, NPT_EINVAL, "bad argument");
  *out = nullptr;
  npt_mg* m = new npt_mg();
  m->cfg = *cfg;
  m->devices.assign(devices, devices + n_devices);
  for (int a = 0; a < n_devices; a++)
    for (int b = a + 1; b < n_devices; b++)
      if (devices[a] == devices[b]) m->use_local = true;   // several contexts on one device (single-GPU tests): no NCCL
  for (int d = 0; d < n_devices; d++) {
    npt_config c = *cfg;
    c.device = devices[d];
    npt_ctx* x = nullptr;
    const int rc = npt_create(&c, &x);
    if (rc) { const std::string why = g_create_error; npt_mg_destroy(m); return mg_fail(nullptr, rc, "device %d: %s", devices[d], why.c_str()); }
    m->ctx.push_back(x);
  }
  m->scratch.resize((size_t)n_devices * 3);
  if (n_devices > 1) {
    if (m->use_local) {
      m->local.members = m->ctx;
      for (int d = 0; d < n_devices; d++) { m->ctx[d]->comm = new NptComm(); m->ctx[d]->comm->local = &m->local; m->ctx[d]->comm_rank = d; m->ctx[d]->comm_world = n_devices; }
    } else {
      if (!nccl().ok) { const std::string why = nccl().why; npt_mg_destroy(m); return mg_fail(nullptr, NPT_ENODEVICE, "NCCL not available: %s", why.c_str()); }
      std::vector<ncclComm_t> comms(n_devices);
      const ncclResult_t r = nccl().CommInitAll(comms.data(), n_devices, m->devices.data());
      if (r != ncclSuccess) { const std::string why = nccl().GetErrorString(r); npt_mg_destroy(m); return mg_fail(nullptr, NPT_ECUDA, "ncclCommInitAll: %s", why.c_str()); }
      for (int d = 0; d < n_devices; d++) { m->ctx[d]->comm = new NptComm(); m->ctx[d]->comm->nc = comms[d]; m->ctx[d]->comm_rank = d; m->ctx[d]->comm_world = n_devices; }
    }
  }
  *out = m;
  return NPT_OK;
}

int npt_mg_set_read_index_base(npt_mg* m, int64_t base) {
  if (!m) return NPT_EINVAL;
  if (base < 0 || base >= (1LL << 40)) return mg_fail(m, NPT_EINVAL, "read index base out of range");
  m->user_base = base; m->next_index = base;
  return NPT_OK;
}

int npt_mg_begin_chrom(npt_mg* m, int32_t chrom_index, const uint8_t* ref_codes, int64_t ref_len, const npt_annotation* annot) {
  if (!m) return NPT_EINVAL;
  m->part.clear(); m->n_batches = 0; m->next_index = m->user_base;
  m->local.reset();
  return mg_parallel(m, [&](int d) { return npt_begin_chrom(m->ctx[d], chrom_index, ref_codes, ref_len, annot); });
}

}  // extern "C"

// Reads [lo_d, lo_{d+1}) of the batch go to device d: contiguous ranges of the submit order with about the same number of
// bases each.  Every device thread rebases its share's offsets (12 bytes per read) and calls npt_submit / npt_submit_packed.
// A packed share must start on an even seq4 byte (two seq4 bytes share one seq2 byte): its first read moves up to the next
// read that does.
static int mg_submit(npt_mg* m, const npt_batch* b, const npt_packed_batch* pb) {
  const int64_t n = b ? b->n : pb->n;
  if (n < 0) return mg_fail(m, NPT_EINVAL, "batch size out of range");
  if (n == 0) return NPT_OK;
  const uint32_t* coff = b ? b->cigar_off : pb->cigar_off;
  const uint64_t* soff = b ? b->seq_off : pb->seq_off;
  const int D = (int)m->ctx.size();
  std::vector<int64_t> lo(D + 1, 0);
  lo[D] = n;
  const uint64_t total = soff[n];
  for (int d = 1; d < D; d++) {
    const uint64_t want = total / D * d;
    lo[d] = std::lower_bound(soff, soff + n, want) - soff;
    lo[d] = std::max(lo[d], lo[d - 1]);
    if (pb) while (lo[d] < n && (soff[lo[d]] & 1)) lo[d]++;
  }
  const int slot = (int)(m->n_batches % 3);
  const int64_t base = m->next_index;
  const int rc = mg_parallel(m, [&](int d) -> int {
    npt_ctx* ctx = m->ctx[d];
    const int64_t a = lo[d], cnt = lo[d + 1] - lo[d];
    if (cnt <= 0) return NPT_OK;
    CK(cudaSetDevice(ctx->device));
    npt_mg::Scratch& s = m->scratch[(size_t)d * 3 + slot];
    if (s.cap < (size_t)cnt + 1) {
      if (s.coff) cudaFreeHost(s.coff);
      if (s.soff) cudaFreeHost(s.soff);
      s.coff = nullptr; s.soff = nullptr; s.cap = 0;
      const size_t cap = (size_t)cnt + 1 + (size_t)cnt / 8;
      CK(cudaHostAlloc(&s.coff, cap * 4, cudaHostAllocDefault));
      CK(cudaHostAlloc(&s.soff, cap * 8, cudaHostAllocDefault));
      s.cap = cap;
    }
    const uint32_t c0 = coff[a], c1 = coff[a + cnt];
    const uint64_t s0 = soff[a], s1 = soff[a + cnt];
    for (int64_t i = 0; i <= cnt; i++) { s.coff[i] = coff[a + i] - c0; s.soff[i] = soff[a + i] - s0; }
    ctx->idx_base = base + a;   // global submit index of the share's first read
    if (b) {
      npt_batch sub = *b;
      sub.n = cnt; sub.pos = b->pos + a; sub.flag = b->flag + a; sub.src = b->src + a;
      sub.cigar_off = s.coff; sub.cigar = b->cigar + c0; sub.seq_off = s.soff; sub.seq4 = b->seq4 + s0;
      return npt_submit(ctx, &sub);
    }
    // the share's part of the wide-element and exception lists, indices rebased to the share
    const uint32_t* w0 = std::lower_bound(pb->wide_idx, pb->wide_idx + pb->n_wide, c0);
    const uint32_t* w1 = std::lower_bound(pb->wide_idx, pb->wide_idx + pb->n_wide, c1);
    const uint64_t* e0 = std::lower_bound(pb->exc_byte, pb->exc_byte + pb->n_exc, s0);
    const uint64_t* e1 = std::lower_bound(pb->exc_byte, pb->exc_byte + pb->n_exc, s1);
    const size_t nw = (size_t)(w1 - w0), ne = (size_t)(e1 - e0);
    if (s.wcap < nw + 1) { if (s.widx) cudaFreeHost(s.widx); s.widx = nullptr; s.wcap = 0; CK(cudaHostAlloc(&s.widx, (nw + 1 + nw / 2) * 4, cudaHostAllocDefault)); s.wcap = nw + 1 + nw / 2; }
    if (s.ecap < ne + 1) { if (s.eidx) cudaFreeHost(s.eidx); s.eidx = nullptr; s.ecap = 0; CK(cudaHostAlloc(&s.eidx, (ne + 1 + ne / 2) * 8, cudaHostAllocDefault)); s.ecap = ne + 1 + ne / 2; }
    for (size_t i = 0; i < nw; i++) s.widx[i] = w0[i] - c0;
    for (size_t i = 0; i < ne; i++) s.eidx[i] = e0[i] - s0;
    npt_packed_batch sub = *pb;
    sub.n = cnt; sub.pos = pb->pos + a; sub.flag = pb->flag + a; sub.src = pb->src + a;
    sub.cigar_off = s.coff; sub.cigar16 = pb->cigar16 + c0;
    sub.n_wide = (int64_t)nw; sub.wide_idx = s.widx; sub.wide_word = pb->wide_word + (w0 - pb->wide_idx);
    sub.seq_off = s.soff; sub.seq2 = pb->seq2 + s0 / 2;   // s0 is even (or the share is the first one)
    sub.n_exc = (int64_t)ne; sub.exc_byte = s.eidx; sub.exc_val = pb->exc_val + (e0 - pb->exc_byte);
    return npt_submit_packed(ctx, &sub);
  });
  if (rc) return rc;
  std::vector<int64_t> cnts(D);
  for (int d = 0; d < D; d++) cnts[d] = lo[d + 1] - lo[d];
  m->part.push_back(cnts);
  m->n_batches++;
  m->next_index += n;
  return NPT_OK;
}

extern "C" {

int npt_mg_submit(npt_mg* m, const npt_batch* b) {
  if (!m || !b) return mg_fail(m, NPT_EINVAL, "null argument");
  if (b->location != NPT_MEM_HOST) return mg_fail(m, NPT_EINVAL, "npt_mg_submit takes host batches (the library splits them over its devices)");
  return mg_submit(m, b, nullptr);
}

int npt_mg_submit_packed(npt_mg* m, const npt_packed_batch* pb) {
  if (!m || !pb) return mg_fail(m, NPT_EINVAL, "null argument");
  if (pb->n_wide < 0 || pb->n_exc < 0) return mg_fail(m, NPT_EINVAL, "packed batch: wide / exception lists inconsistent");
  return mg_submit(m, nullptr, pb);
}

int npt_mg_wait(npt_mg* m) {
  if (!m) return NPT_EINVAL;
  return mg_parallel(m, [&](int d) { return npt_wait(m->ctx[d]); });
}

int npt_mg_end_chrom(npt_mg* m) {
  if (!m) return NPT_EINVAL;
  return mg_parallel(m, [&](int d) { return end_chrom_all(m->ctx[d]); });
}

// Tables and coverage are global on every device after npt_mg_end_chrom: they come from device 0.  Per-read arrays come
// from every device and are put back into submit order (batch by batch, device by device).
int npt_mg_fetch_results(npt_mg* m, int what, npt_results* out) {
  if (!m || !out) return mg_fail(m, NPT_EINVAL, "null argument");
  const int D = (int)m->ctx.size();
  std::vector<npt_results> rs(D);
  int rc = mg_parallel(m, [&](int d) { return npt_fetch_results(m->ctx[d], d == 0 ? what : (what & NPT_FETCH_READS), &rs[d]); });
  if (rc) return rc;
  npt_results R = rs[0];
  int64_t n_total = 0, slow = 0;
  for (int d = 0; d < D; d++) { n_total += rs[d].n_reads; slow += rs[d].n_slow_reads; }
  R.n_reads = n_total; R.n_slow_reads = slow;
  if ((what & NPT_FETCH_READS) && D > 1) {
    const size_t nn = (size_t)std::max<int64_t>(n_total, 1);
    for (auto& v : m->r_i32) v.resize(nn);
    m->r_flags.resize(nn); m->r_boff.resize(nn + 1);
    uint64_t nb_total = 0;
    for (int d = 0; d < D; d++) nb_total += rs[d].n_reads ? rs[d].break_off[rs[d].n_reads] : 0;
    m->r_breaks.resize((size_t)std::max<uint64_t>(nb_total, 1));
    std::vector<int64_t> cur(D, 0);
    int64_t o = 0;
    uint64_t bo = 0;
    for (auto& cnts : m->part)
      for (int d = 0; d < D; d++) {
        const int64_t c = cnts[d], a = cur[d];
        if (!c) continue;
        const npt_results& r = rs[d];
        const int32_t* src8[8] = {r.cluster_id, r.sub_id, r.start_pos, r.end_pos, r.read_st, r.read_en, r.maps_sum, r.err_sum};
        for (int k = 0; k < 8; k++) memcpy(m->r_i32[k].data() + o, src8[k] + a, (size_t)c * 4);
        memcpy(m->r_flags.data() + o, r.read_flags + a, (size_t)c * 4);
        const uint64_t b0 = r.break_off[a], b1 = r.break_off[a + c];
        for (int64_t i = 0; i < c; i++) m->r_boff[o + i] = r.break_off[a + i] - b0 + bo;
        if (b1 > b0) memcpy(m->r_breaks.data() + bo, r.breaks + b0, (size_t)(b1 - b0) * 4);
        bo += b1 - b0; o += c; cur[d] += c;
      }
    m->r_boff[n_total] = bo;
    R.cluster_id = m->r_i32[0].data(); R.sub_id = m->r_i32[1].data(); R.start_pos = m->r_i32[2].data(); R.end_pos = m->r_i32[3].data();
    R.read_st = m->r_i32[4].data(); R.read_en = m->r_i32[5].data(); R.maps_sum = m->r_i32[6].data(); R.err_sum = m->r_i32[7].data();
    R.read_flags = m->r_flags.data(); R.break_off = m->r_boff.data(); R.breaks = m->r_breaks.data();
  }
  m->res = R;
  *out = R;
  return NPT_OK;
}

int npt_mg_release_results(npt_mg* m) {
  if (!m) return NPT_EINVAL;
  return mg_parallel(m, [&](int d) { return npt_release_results(m->ctx[d]); });
}

int npt_mg_last_step_ms(npt_mg* m, float* step_ms, float* key_exchange_ms, float* reduce_ms) {
  if (!m) return NPT_EINVAL;
  float s = 0, k = 0, r = 0;
  for (auto* c : m->ctx) { s = std::max(s, c->step_ms); k = std::max(k, c->kx_ms); r = std::max(r, c->reduce_ms); }
  if (step_ms) *step_ms = s;
  if (key_exchange_ms) *key_exchange_ms = k;
  if (reduce_ms) *reduce_ms = r;
  return NPT_OK;
}

}  // extern "C"
