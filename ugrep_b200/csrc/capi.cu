// capi.cu — the C ABI declared in include/ugrep_b200.h: pattern upload, scanner scratch,
// and the scan entry points.  Host C++ only calls CUDA through this file and scan_kernels.cu.
// There is no CPU scan path: without a usable CUDA device every entry point fails.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/ugrep_b200.h"
#include "device_pattern.cuh"
#include "pattern_host.hpp"
#include "scan_kernels.hpp"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg)
{
  g_err = msg;
  return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return UGX_E_CUDA;
}

#define CU(call)                         \
  do                                     \
  {                                      \
    cudaError_t e_ = (call);             \
    if (e_ != cudaSuccess)               \
      return cuda_fail(e_, #call);       \
  } while (0)

const int k_word_ranges[] = {
#include "word_ranges.inc"
};

template <typename T>
int upload(T*& dev, const void* host, size_t bytes)
{
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
  if (e != cudaSuccess)
    return cuda_fail(e, "cudaMalloc");
  if (bytes)
  {
    e = cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
    {
      cudaFree(p);
      return cuda_fail(e, "cudaMemcpy");
    }
  }
  dev = static_cast<T*>(p);
  return UGX_OK;
}

void set256(uint32_t* set, const uint8_t* bytes32)
{
  for (int i = 0; i < 8; ++i)
    set[i] = static_cast<uint32_t>(bytes32[4 * i]) | (static_cast<uint32_t>(bytes32[4 * i + 1]) << 8) |
             (static_cast<uint32_t>(bytes32[4 * i + 2]) << 16) | (static_cast<uint32_t>(bytes32[4 * i + 3]) << 24);
}

} // namespace

struct ugx_pattern {
  ugx::HostDfa dfa;
  ugx_prefilter pf;
  uint32_t flags = 0;
  int adv = 0;
  bool never = false; // the prefilter tables admit no candidate at all (config 3): every scan finds nothing
  int device = 0;
  uint32_t nop = 0;
  ugx::DevPattern dev;
  std::vector<void*> allocs;
  ~ugx_pattern()
  {
    for (void* p : allocs)
      cudaFree(p);
  }
};

struct ugx_scanner {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // scratch
  uint64_t* tile_matches = nullptr;
  uint64_t* tile_newlines = nullptr;
  uint64_t tiles_cap = 0;
  uint32_t* strip_counts = nullptr;
  uint64_t strips_cap = 0;
  unsigned long long* totals = nullptr; // device [4]
  unsigned long long* h_totals = nullptr; // pinned [4]
  ugx_match* records = nullptr;
  uint64_t records_cap = 0;
  uint64_t records_n = 0;
  uint8_t* stage = nullptr; // device copy of a host buffer
  bool force_generic = false; // tests: always take the generic line-scan kernel
  bool legacy_any = false;    // tests / A-B timing: the tile-synchronous count_lines_any kernel instead of the streaming one
  bool count_newlines = false; // the streaming count also counts newlines
  bool match_lines = false;      // counting takes match_lines_kernel (position-parallel attempts) instead of scan_lines_kernel
  bool two_pass_records = false; // records by count pass + emit pass (A/B against the single-pass staging form)
  uint64_t* tile_base = nullptr;  // staging base of every tile's records
  uint64_t tile_base_cap = 0;
  ugx_match* rec_stage = nullptr; // records in tile completion order
  uint64_t rec_stage_cap = 0;
  bool no_pipeline = false;    // host buffers: one copy, then the scan (A/B timing of the overlapped path)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr;
  bool stream_dfa = false;     // DFA patterns take the streaming count too (default: the tile-synchronous kernel)
  // streaming count scratch
  uint8_t* region_sum = nullptr;
  uint64_t region_cap = 0;
  unsigned long long* partials = nullptr; // [2 * STREAM_MAX_GRID]
  unsigned long long* ticket = nullptr;   // {ticket (u64), done (u32)}
  uint64_t stage_cap = 0;
  // span scan scratch
  uint64_t* span_regions = nullptr; // [5 * regions]
  uint64_t span_regions_cap = 0;
  // pageable host buffers: feeder threads copy chunks into pinned slots, each slot goes to the device on its thread's stream
  std::vector<uint8_t*> feed_slots;      // [2 * feed_threads] pinned, FEED_CHUNK bytes each
  std::vector<cudaStream_t> feed_streams; // [feed_threads]
  std::vector<cudaEvent_t> feed_events;   // [2 * feed_threads] slot free again
  bool no_feeder = false;         // pageable buffers take one plain cudaMemcpyAsync (A/B timing)
  uint16_t* span_sel = nullptr;   // records: selected match starts per 16-byte chunk
  uint64_t span_sel_cap = 0;
  uint64_t* batch = nullptr;      // ugx_count_batch: file table, counters, tile table
  uint64_t batch_cap = 0;
  bool no_span = false;           // counting matches / records take the line-at-a-time kernels (A/B timing, tests)
  bool no_cover = false;          // the streaming count evaluates the candidate predicate even where it is proven implied
};

extern "C" {

const char* ugx_last_error(void) { return g_err.c_str(); }
int ugx_abi_version(void) { return UGX_ABI_VERSION; }

static int pattern_create_impl(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, int device,
                               ugx_pattern** out);

// no C++ exception may cross the C ABI: allocation failures inside the export become status codes
int ugx_pattern_create(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, int device,
                       ugx_pattern** out)
{
  try
  {
    return pattern_create_impl(opc, nop, pf, matcher_flags, device, out);
  }
  catch (const std::bad_alloc&)
  {
    return fail(UGX_E_NOMEM, "out of host memory");
  }
  catch (const std::exception& ex)
  {
    return fail(UGX_E_INVALID, std::string("ugx_pattern_create: ") + ex.what());
  }
}

static int pattern_create_impl(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, int device,
                               ugx_pattern** out)
{
  if (opc == nullptr || nop == 0 || pf == nullptr || out == nullptr)
    return fail(UGX_E_INVALID, "ugx_pattern_create: null argument");
  std::unique_ptr<ugx_pattern> holder(new (std::nothrow) ugx_pattern()); // freed on every early return and on a throw
  ugx_pattern* p = holder.get();
  if (p == nullptr)
    return fail(UGX_E_NOMEM, "out of host memory");
  std::string err;
  int rc = ugx::flatten_dfa(opc, nop, p->dfa, err);
  if (rc == UGX_OK)
    rc = ugx::check_scope(p->dfa, *pf, matcher_flags, err);
  if (rc != UGX_OK)
  {
    return fail(rc, err);
  }
  p->pf = *pf;
  p->flags = matcher_flags;
  p->adv = ugx::select_advance(*pf, matcher_flags);
  p->never = ugx::prefilter_never_fires(*pf, p->adv);
  p->device = device;
  p->nop = nop;
  cudaError_t ce = cudaSetDevice(device);
  if (ce != cudaSuccess)
  {
    return cuda_fail(ce, "cudaSetDevice (the scan path needs a CUDA device; there is no CPU fallback)");
  }
  ugx::DevPattern& d = p->dev;
  memset(&d, 0, sizeof(d));
  d.adv = p->adv;
  d.len = pf->len;
  d.min = pf->min;
  d.pin = pf->pin;
  d.lcp = pf->lcp;
  d.lcs = pf->lcs;
  d.one = pf->one;
  d.bol = pf->bol;
  d.lbk = pf->lbk;
  d.lbm = pf->lbm;
  d.flags = matcher_flags;
  d.nstates = p->dfa.nstates;
  d.ncls = p->dfa.ncls;
  d.has_meta = p->dfa.has_meta;
  d.to_start = p->dfa.to_start;
  d.nop = nop;
  d.table_bytes = p->dfa.table_bytes();
  d.first_acc = p->dfa.first_acc;
  d.first_leaf = p->dfa.first_leaf;
  d.acc0 = p->dfa.accept[0] != 0;
  ugx::plan_filter(*pf, p->adv, d.plan);
  d.covers = ugx::prefilter_covers_matches(p->dfa, *pf, p->adv, matcher_flags) ? 1u : 0u;
  d.n_word_ranges = sizeof(k_word_ranges) / sizeof(int) / 2;
  memcpy(d.chr, pf->chr, 256);
  set256(d.cbk, pf->cbk);
  set256(d.fst, pf->fst);
  if (pf->len == 0 && pf->pin >= 1 && pf->pin <= 16)
  {
    for (uint32_t i = 0; i < pf->pin; ++i)
    {
      uint32_t a = pf->chr[i], b = pf->chr[pf->pin + i];
      d.pin_a[a >> 5] |= 1u << (a & 31);
      d.pin_b[b >> 5] |= 1u << (b & 31);
    }
  }
  // accept words: bit 31 marks a state without outgoing byte edges (the interpreter halts there before reading)
  std::vector<uint32_t> acc(p->dfa.nstates);
  for (uint32_t s = 0; s < p->dfa.nstates; ++s)
  {
    bool any = false;
    for (uint32_t c = 0; c < p->dfa.ncls && !any; ++c)
      any = p->dfa.next[static_cast<size_t>(s) * p->dfa.ncls + c] != ugx::DEAD;
    acc[s] = (p->dfa.accept[s] & 0x7fffffffu) | (any ? 0u : 0x80000000u);
  }
  std::vector<uint32_t> opcp(opc, opc + nop);
  opcp.push_back(ugx::OP_HALT);
  opcp.push_back(ugx::OP_HALT);
  std::vector<uint16_t> next = p->dfa.next;
  next.resize((next.size() + 7) / 8 * 8, ugx::DEAD); // whole uint4s for the staging loop
  ugx::Viability via;
  ugx::build_viability(p->dfa, via);
  d.via_k = via.k;
  d.via_stride = via.stride;
  std::vector<uint32_t> via_t(512);
  for (int b = 0; b < 256; ++b)
  {
    via_t[2 * b] = via.t01[b];
    via_t[2 * b + 1] = via.t23[b];
  }
  if (via.k != 0 && via.bits.size() * 32 <= 8192)
  {
    // a small table is stored one byte per entry: the device then needs one load per position instead of load, shift, mask
    std::vector<uint32_t> bytes(via.bits.size() * 8, 0);
    for (size_t i = 0; i < via.bits.size() * 32; ++i)
      if ((via.bits[i >> 5] >> (i & 31)) & 1u)
        bytes[i >> 2] |= 1u << (8 * (i & 3));
    via.bits.swap(bytes);
    d.via_bytes = 1;
  }
  via.bits.resize((via.bits.size() + 3) / 4 * 4 + 4, 0); // whole uint4s for the staging copy
  via.pair.resize((via.pair.size() + 15) / 16 * 16 + 16, 0);
  d.via_words = static_cast<uint32_t>(via.bits.size());
  d.via_pair_bytes = static_cast<uint32_t>(via.pair.size());
  uint32_t* d_vids = nullptr;
  uint32_t* d_vbits = nullptr;
  uint8_t* d_vpair = nullptr;
  uint8_t* d_cls = nullptr;
  uint16_t* d_next = nullptr;
  uint32_t* d_acc = nullptr;
  uint32_t* d_opc = nullptr;
  uint8_t* d_pred = nullptr;
  uint8_t* d_tap = nullptr;
  int* d_words = nullptr;
  rc = upload(d_cls, p->dfa.cls, 256);
  if (rc == UGX_OK) { p->allocs.push_back(d_cls); rc = upload(d_next, next.data(), next.size() * 2); }
  if (rc == UGX_OK) { p->allocs.push_back(d_next); rc = upload(d_acc, acc.data(), acc.size() * 4); }
  if (rc == UGX_OK) { p->allocs.push_back(d_acc); rc = upload(d_opc, opcp.data(), opcp.size() * 4); }
  if (rc == UGX_OK) { p->allocs.push_back(d_opc); rc = upload(d_pred, pf->min < 4 ? pf->pma : pf->pmh, UGX_HASH); }
  if (rc == UGX_OK) { p->allocs.push_back(d_pred); rc = upload(d_tap, pf->tap, UGX_BTAP); }
  if (rc == UGX_OK) { p->allocs.push_back(d_tap); rc = upload(d_words, k_word_ranges, sizeof(k_word_ranges)); }
  if (rc == UGX_OK) { p->allocs.push_back(d_words); rc = upload(d_vids, via_t.data(), via_t.size() * 4); }
  if (rc == UGX_OK) { p->allocs.push_back(d_vids); rc = upload(d_vbits, via.bits.data(), via.bits.size() * 4); }
  if (rc == UGX_OK) { p->allocs.push_back(d_vbits); rc = upload(d_vpair, via.pair.data(), via.pair.size()); }
  if (rc == UGX_OK) p->allocs.push_back(d_vpair);
  if (rc != UGX_OK)
  {
    return rc;
  }
  d.cls = d_cls;
  d.next = d_next;
  d.accept = d_acc;
  d.opc = d_opc;
  d.pred = d_pred;
  d.tap = d_tap;
  d.word_ranges = d_words;
  d.via_ids = d_vids;
  d.via_bits = d_vbits;
  d.via_pair = d_vpair;
  *out = holder.release();
  return UGX_OK;
}

static int plan_describe_impl(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, ugx_plan_info* out);

int ugx_plan_describe(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, ugx_plan_info* out)
{
  try
  {
    return plan_describe_impl(opc, nop, pf, matcher_flags, out);
  }
  catch (const std::bad_alloc&)
  {
    return fail(UGX_E_NOMEM, "out of host memory");
  }
  catch (const std::exception& ex)
  {
    return fail(UGX_E_INVALID, std::string("ugx_plan_describe: ") + ex.what());
  }
}

static int plan_describe_impl(const uint32_t* opc, uint32_t nop, const ugx_prefilter* pf, uint32_t matcher_flags, ugx_plan_info* out)
{
  if (opc == nullptr || nop == 0 || pf == nullptr || out == nullptr)
    return fail(UGX_E_INVALID, "ugx_plan_describe: null argument");
  ugx::HostDfa dfa;
  std::string err;
  int rc = ugx::flatten_dfa(opc, nop, dfa, err);
  if (rc != UGX_OK)
    return fail(rc, err);
  memset(out, 0, sizeof(*out));
  out->states = dfa.nstates;
  out->classes = dfa.ncls;
  out->table_bytes = dfa.table_bytes();
  out->first_acc = dfa.first_acc;
  out->first_leaf = dfa.first_leaf;
  out->max_match_len = dfa.max_match_len;
  out->has_meta = dfa.has_meta;
  out->newline_live = dfa.newline_live;
  const int adv = ugx::select_advance(*pf, matcher_flags);
  out->advance = adv;
  out->covers = ugx::prefilter_covers_matches(dfa, *pf, adv, matcher_flags) ? 1u : 0u;
  ugx::FilterPlan plan;
  ugx::plan_filter(*pf, adv, plan);
  out->kind = plan.kind;
  out->nterms = plan.nterms;
  for (int i = 0; i < 3; ++i)
    out->t_off[i] = plan.t_off[i];
  for (int i = 0; i < 2; ++i)
  {
    out->a_off[i] = plan.a_off[i];
    out->a_chr[i] = plan.a_chr[i] & 0xffu;
  }
  out->h4_terms = plan.h4_terms;
  out->h4_shift = plan.h4_shift;
  out->pm2 = plan.pm2;
  out->pm2_shift = plan.pm2_shift;
  memcpy(out->lut, plan.lut, sizeof(out->lut));
  rc = ugx::check_scope(dfa, *pf, matcher_flags, err);
  if (rc != UGX_OK)
    return fail(rc, err);
  return UGX_OK;
}

int ugx_viability_describe(const uint32_t* opc, uint32_t nop, uint32_t* k, uint32_t* stride, uint32_t* t01, uint32_t* t23,
                           uint8_t* pair, uint32_t cap_pair, uint32_t* npair, uint32_t* bits, uint32_t cap_words,
                           uint32_t* words)
{
  if (opc == nullptr || nop == 0 || k == nullptr || stride == nullptr || t01 == nullptr || t23 == nullptr ||
      npair == nullptr || words == nullptr)
    return fail(UGX_E_INVALID, "ugx_viability_describe: null argument");
  try
  {
    ugx::HostDfa dfa;
    std::string err;
    const int rc = ugx::flatten_dfa(opc, nop, dfa, err);
    if (rc != UGX_OK)
      return fail(rc, err);
    ugx::Viability v;
    ugx::build_viability(dfa, v);
    *k = v.k;
    *stride = v.stride;
    memcpy(t01, v.t01, sizeof(v.t01));
    memcpy(t23, v.t23, sizeof(v.t23));
    *npair = static_cast<uint32_t>(v.pair.size());
    *words = static_cast<uint32_t>(v.bits.size());
    if (v.bits.size() > cap_words || v.pair.size() > cap_pair)
      return fail(UGX_E_OVERFLOW, "viability table larger than the caller's buffer");
    if (!v.pair.empty() && pair != nullptr)
      memcpy(pair, v.pair.data(), v.pair.size());
    if (!v.bits.empty() && bits != nullptr)
      memcpy(bits, v.bits.data(), v.bits.size() * 4);
    return UGX_OK;
  }
  catch (const std::exception&)
  {
    return fail(UGX_E_NOMEM, "out of host memory");
  }
}

int ugx_pattern_load(const char* path, int device, ugx_pattern** out)
{
  FILE* f = fopen(path, "rb");
  if (f == nullptr)
    return fail(UGX_E_IO, std::string("cannot open ") + path);
  ugx_file_header h;
  ugx_prefilter pf;
  std::vector<uint32_t> opc;
  bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, UGX_FILE_MAGIC, 8) == 0 &&
            h.prefilter_size == sizeof(pf) && fread(&pf, sizeof(pf), 1, f) == 1;
  if (ok)
  {
    // the opcode count comes from the file: it must fit in what is left of the file before anything is allocated
    const long here = ftell(f);
    ok = here >= 0 && fseek(f, 0, SEEK_END) == 0;
    const long size = ok ? ftell(f) : -1;
    ok = ok && size >= here && h.nop > 0 && static_cast<uint64_t>(h.nop) * 4 <= static_cast<uint64_t>(size - here) &&
         fseek(f, here, SEEK_SET) == 0;
  }
  if (ok)
  {
    try
    {
      opc.resize(h.nop);
    }
    catch (const std::exception&)
    {
      fclose(f);
      return fail(UGX_E_NOMEM, "out of host memory");
    }
    ok = fread(opc.data(), 4, h.nop, f) == h.nop;
  }
  fclose(f);
  if (!ok)
    return fail(UGX_E_IO, std::string("not a UGXP pattern file: ") + path);
  return ugx_pattern_create(opc.data(), h.nop, &pf, h.matcher_flags, device, out);
}

int ugx_pattern_info_get(const ugx_pattern* p, ugx_pattern_info* info)
{
  if (p == nullptr || info == nullptr)
    return fail(UGX_E_INVALID, "null argument");
  info->nop = p->nop;
  info->states = p->dfa.nstates;
  info->classes = p->dfa.ncls;
  info->table_bytes = p->dfa.table_bytes();
  info->table_in_smem = !p->dfa.has_meta && p->dfa.table_bytes() <= ugx::SCAN_MAX_SMEM_TABLE;
  info->advance = p->adv;
  info->has_meta = p->dfa.has_meta;
  info->lookback = p->pf.lbk != 0;
  return UGX_OK;
}

void ugx_pattern_destroy(ugx_pattern* p)
{
  if (p != nullptr)
  {
    cudaSetDevice(p->device);
    delete p;
  }
}

int ugx_scanner_create(int device, void* stream, ugx_scanner** out)
{
  if (out == nullptr)
    return fail(UGX_E_INVALID, "null argument");
  CU(cudaSetDevice(device));
  ugx_scanner* s = new (std::nothrow) ugx_scanner();
  if (s == nullptr)
    return fail(UGX_E_NOMEM, "out of host memory");
  s->device = device;
  s->stream = static_cast<cudaStream_t>(stream);
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e == cudaSuccess)
    e = cudaEventCreate(&s->ev0);
  if (e == cudaSuccess)
    e = cudaEventCreate(&s->ev1);
  if (e == cudaSuccess)
    e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess)
    e = cudaEventCreateWithFlags(&s->ev_copy, cudaEventDisableTiming);
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(&s->totals), 8 * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaMallocHost(reinterpret_cast<void**>(&s->h_totals), 8 * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(&s->partials), 2 * ugx::STREAM_MAX_GRID * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaMalloc(reinterpret_cast<void**>(&s->ticket), 2 * sizeof(unsigned long long));
  if (e == cudaSuccess)
    e = cudaMemset(s->ticket, 0, 2 * sizeof(unsigned long long));
  if (e != cudaSuccess)
  {
    ugx_scanner_destroy(s);
    return cuda_fail(e, "ugx_scanner_create");
  }
  s->sm_count = prop.multiProcessorCount;
  *out = s;
  return UGX_OK;
}

void ugx_scanner_destroy(ugx_scanner* s)
{
  if (s == nullptr)
    return;
  cudaSetDevice(s->device);
  cudaFree(s->tile_matches);
  cudaFree(s->tile_newlines);
  cudaFree(s->strip_counts);
  cudaFree(s->totals);
  cudaFreeHost(s->h_totals);
  cudaFree(s->region_sum);
  cudaFree(s->span_regions);
  cudaFree(s->span_sel);
  for (uint8_t* p : s->feed_slots)
    cudaFreeHost(p);
  for (cudaEvent_t e : s->feed_events)
    cudaEventDestroy(e);
  for (cudaStream_t st : s->feed_streams)
    cudaStreamDestroy(st);
  cudaFree(s->batch);
  cudaFree(s->tile_base);
  cudaFree(s->rec_stage);
  cudaFree(s->partials);
  cudaFree(s->ticket);
  cudaFree(s->records);
  cudaFree(s->stage);
  if (s->ev0)
    cudaEventDestroy(s->ev0);
  if (s->ev1)
    cudaEventDestroy(s->ev1);
  if (s->ev_copy)
    cudaEventDestroy(s->ev_copy);
  if (s->copy_stream)
    cudaStreamDestroy(s->copy_stream);
  delete s;
}

} // extern "C"

namespace {

template <typename T>
int ensure(T*& ptr, uint64_t& cap, uint64_t need)
{
  if (need <= cap)
    return UGX_OK;
  if (ptr != nullptr)
    cudaFree(ptr);
  ptr = nullptr;
  cap = 0;
  uint64_t want = need + need / 8 + 64;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want * sizeof(T));
  if (e != cudaSuccess)
  {
    g_err = std::string("cudaMalloc scratch: ") + cudaGetErrorString(e);
    return e == cudaErrorMemoryAllocation ? UGX_E_NOMEM : UGX_E_CUDA;
  }
  ptr = static_cast<T*>(p);
  cap = want;
  return UGX_OK;
}

constexpr uint64_t FEED_CHUNK = 8ull << 20; // bytes per pinned slot of the pageable feeder
constexpr uint64_t FEED_MIN = 32ull << 20;  // smaller pageable buffers take one plain copy

// Host-to-device copy of a PAGEABLE buffer (what an mmap'ing caller such as ugrep hands over).  cudaMemcpyAsync from
// pageable memory goes through the driver's own staging at ~10 GB/s on this box; here K host threads copy 8 MiB chunks
// into pinned slots (two per thread) and send each slot on the thread's own stream, so the memcpy's of some chunks
// overlap the DMA of others.  The scan stream waits on every thread's last copy.
int feed_pageable(ugx_scanner* s, uint8_t* dst, const uint8_t* src, uint64_t n)
{
  if (s->feed_streams.empty())
  {
    // half the host threads, shared between the ranks of this box when there are several (torchrun's LOCAL_WORLD_SIZE)
    unsigned hw = std::thread::hardware_concurrency();
    unsigned ranks = 1;
    if (const char* lws = getenv("LOCAL_WORLD_SIZE"))
      ranks = static_cast<unsigned>(atoi(lws)) > 0 ? static_cast<unsigned>(atoi(lws)) : 1;
    unsigned k = hw / (2 * ranks);
    if (const char* env = getenv("UGX_FEED_THREADS"))
      k = static_cast<unsigned>(atoi(env));
    if (k < 2)
      k = 2;
    if (k > 8)
      k = 8;
    // all or nothing: a feeder that could not get its streams / pinned slots leaves no half-built state behind
    cudaError_t e = cudaSuccess;
    for (unsigned i = 0; i < k && e == cudaSuccess; ++i)
    {
      cudaStream_t st = nullptr;
      e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      if (e != cudaSuccess)
        break;
      s->feed_streams.push_back(st);
      for (int j = 0; j < 2 && e == cudaSuccess; ++j)
      {
        void* p = nullptr;
        e = cudaHostAlloc(&p, FEED_CHUNK, cudaHostAllocDefault);
        if (e != cudaSuccess)
          break;
        s->feed_slots.push_back(static_cast<uint8_t*>(p));
        cudaEvent_t ev = nullptr;
        e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        if (e == cudaSuccess)
          s->feed_events.push_back(ev);
      }
    }
    if (e != cudaSuccess)
    {
      for (uint8_t* p : s->feed_slots)
        cudaFreeHost(p);
      for (cudaEvent_t ev : s->feed_events)
        cudaEventDestroy(ev);
      for (cudaStream_t st : s->feed_streams)
        cudaStreamDestroy(st);
      s->feed_slots.clear();
      s->feed_events.clear();
      s->feed_streams.clear();
      return cuda_fail(e, "pageable feeder setup");
    }
  }
  const unsigned K = static_cast<unsigned>(s->feed_streams.size());
  const uint64_t nchunks = (n + FEED_CHUNK - 1) / FEED_CHUNK;
  std::vector<cudaError_t> errs(K, cudaSuccess);
  std::vector<std::thread> workers;
  for (unsigned w = 0; w < K; ++w)
    workers.emplace_back([&, w]() {
      cudaError_t e = cudaSetDevice(s->device);
      uint64_t round = 0;
      for (uint64_t c = w; c < nchunks && e == cudaSuccess; c += K, ++round)
      {
        const unsigned slot = 2 * w + static_cast<unsigned>(round & 1);
        if (round >= 2)
          e = cudaEventSynchronize(s->feed_events[slot]); // the slot's previous copy has left it
        if (e != cudaSuccess)
          break;
        const uint64_t off = c * FEED_CHUNK;
        const uint64_t len = n - off < FEED_CHUNK ? n - off : FEED_CHUNK;
        memcpy(s->feed_slots[slot], src + off, len);
        e = cudaMemcpyAsync(dst + off, s->feed_slots[slot], len, cudaMemcpyHostToDevice, s->feed_streams[w]);
        if (e == cudaSuccess)
          e = cudaEventRecord(s->feed_events[slot], s->feed_streams[w]);
      }
      errs[w] = e;
    });
  for (auto& t : workers)
    t.join();
  for (unsigned w = 0; w < K; ++w)
    if (errs[w] != cudaSuccess)
      return cuda_fail(errs[w], "pageable feeder");
  // the scan stream continues after the last copy of every feeder stream
  for (unsigned w = 0; w < K; ++w)
  {
    CU(cudaEventRecord(s->ev_copy, s->feed_streams[w]));
    CU(cudaStreamWaitEvent(s->stream, s->ev_copy, 0));
  }
  return UGX_OK;
}

// make `buf` visible to the device: device pointers are used in place, host pointers are staged
int resolve(ugx_scanner* s, const void* buf, uint64_t n, const uint8_t** dev, uint64_t* h2d)
{
  *h2d = 0;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, buf);
  if (e == cudaSuccess && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged))
  {
    if ((reinterpret_cast<uintptr_t>(buf) & 15) == 0)
    {
      *dev = static_cast<const uint8_t*>(buf);
      return UGX_OK;
    }
    // the kernels use 16-byte vector loads: a misaligned device buffer is copied to aligned scratch
    int rc = ensure(s->stage, s->stage_cap, n + 16);
    if (rc != UGX_OK)
      return rc;
    CU(cudaMemcpyAsync(s->stage, buf, n, cudaMemcpyDeviceToDevice, s->stream));
    *dev = s->stage;
    return UGX_OK;
  }
  if (e != cudaSuccess)
    cudaGetLastError(); // plain host memory on older drivers reports an error: clear it
  int rc = ensure(s->stage, s->stage_cap, n + 16);
  if (rc != UGX_OK)
    return rc;
  const bool pageable = e != cudaSuccess || at.type == cudaMemoryTypeUnregistered;
  if (pageable && n >= FEED_MIN && !s->no_feeder)
  {
    // earlier work on the scan stream may still read the staging buffer
    CU(cudaStreamSynchronize(s->stream));
    rc = feed_pageable(s, s->stage, static_cast<const uint8_t*>(buf), n);
    if (rc != UGX_OK)
      return rc;
  }
  else
    CU(cudaMemcpyAsync(s->stage, buf, n, cudaMemcpyHostToDevice, s->stream));
  *dev = s->stage;
  *h2d = n;
  return UGX_OK;
}

constexpr uint64_t PIPE_CHUNK = 32ull << 20; // bytes per host-to-device chunk of the pipelined path (region-aligned)

bool is_host_pointer(const void* buf)
{
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, buf);
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    return true;
  }
  return at.type != cudaMemoryTypeDevice && at.type != cudaMemoryTypeManaged;
}

// page-locked host memory (cudaHostAlloc / cudaHostRegister): DMA reads it directly
bool is_pinned_host_pointer(const void* buf)
{
  cudaPointerAttributes at;
  const cudaError_t e = cudaPointerGetAttributes(&at, buf);
  if (e != cudaSuccess)
  {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

// `ugrep -c -o` / `ugrep -o -n -b` through the span kernels (span_scan.cu).  *valid = false: the spans could not vouch
// for their result on this buffer (a match longer than a window across a region start in a line without newlines, an
// attempt that failed at the very end of the buffer, a match of 64 KiB or more) — the caller then takes the
// line-at-a-time kernels.
int scan_spans(ugx_scanner* s, const ugx_pattern* p, const uint8_t* dbuf, uint64_t n, bool want_records, uint64_t base_offset,
               uint64_t base_line, ugx_totals* tt, bool* valid)
{
  *valid = false;
  const uint64_t nreg = ugx::stream_regions(n);
  int rc = ensure(s->span_regions, s->span_regions_cap, 5 * nreg + 8);
  if (rc == UGX_OK && want_records)
    rc = ensure(s->span_sel, s->span_sel_cap, (n + 15) / 16 + 64);
  if (rc != UGX_OK)
    return rc;
  ugx::SpanArgs a;
  memset(&a, 0, sizeof(a));
  a.reg_matches = s->span_regions;
  a.reg_newlines = s->span_regions + nreg;
  a.reg_emain = s->span_regions + 2 * nreg;
  a.reg_elast = s->span_regions + 3 * nreg;
  a.reg_v = s->span_regions + 4 * nreg;
  a.sel_bits = want_records ? s->span_sel : nullptr;
  a.base_offset = base_offset;
  a.base_line = base_line;
  a.no_cover = s->no_cover ? 1u : 0u;
  a.tail = reinterpret_cast<const uint64_t*>(s->totals + 5);
  a.flags = reinterpret_cast<unsigned int*>(s->totals + 6);
  CU(cudaMemsetAsync(s->totals + 6, 0, sizeof(unsigned long long), s->stream));
  CU(ugx::launch_last_line(dbuf, n, reinterpret_cast<uint64_t*>(s->totals + 5), s->stream));
  CU(ugx::launch_span_scan(p->dev, dbuf, n, a, false, s->sm_count, s->stream));
  CU(ugx::launch_span_final(p->dev, dbuf, n, a, false, s->totals, s->stream));
  CU(cudaMemcpyAsync(s->h_totals, s->totals, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  tt->launches += 3;
  if (s->h_totals[2] != 0)
    return UGX_OK; // not valid: nothing of this attempt is used
  const uint64_t nrec = s->h_totals[0];
  if (want_records)
  {
    rc = ensure(s->records, s->records_cap, nrec);
    if (rc != UGX_OK)
      return rc;
    if (nrec > 0)
    {
      a.out = s->records;
      a.out_cap = nrec;
      CU(ugx::launch_span_scan(p->dev, dbuf, n, a, true, s->sm_count, s->stream));
      CU(ugx::launch_span_final(p->dev, dbuf, n, a, true, s->totals, s->stream));
      tt->launches += 2;
    }
    s->records_n = nrec;
  }
  tt->matches = nrec;
  tt->newlines = s->h_totals[1];
  tt->flags |= UGX_TOT_NEWLINES;
  tt->kernel = UGX_K_SPAN;
  *valid = true;
  return UGX_OK;
}

int scan_common(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t n, int mode, bool want_records,
                uint64_t base_offset, uint64_t base_line, const ugx_match** dev_out, uint64_t* n_out, ugx_totals* totals)
{
  if (s == nullptr || p == nullptr || (buf == nullptr && n != 0))
    return fail(UGX_E_INVALID, "null argument");
  if (s->device != p->device)
    return fail(UGX_E_INVALID, "pattern and scanner live on different devices");
  ugx_totals tt;
  memset(&tt, 0, sizeof(tt));
  if (n_out)
    *n_out = 0;
  if (dev_out)
    *dev_out = nullptr;
  if (n == 0)
  {
    if (totals)
      *totals = tt;
    return UGX_OK;
  }
  CU(cudaSetDevice(s->device));
  const uint8_t* dbuf = nullptr;
  uint64_t h2d = 0;
  const bool stream_route = mode == 0 && !want_records && !s->force_generic && !s->legacy_any &&
                            ugx::count_lines_stream_eligible(p->dev) &&
                            (s->stream_dfa || ugx::count_lines_literal_eligible(p->dev) || p->dev.plan.h4_terms >= 1 || p->dev.plan.pm2 != 0 ||
                             (p->dev.plan.kind == ugx::FK_LUT && p->dev.plan.nterms >= 1 && p->dev.one == 0));
  // a host buffer on the streaming route is copied chunk by chunk, overlapped with the scan, when every read of a
  // scanned position stays within one region of it: pure literals, and DFAs whose longest match is bounded
  const bool bounded = ugx::count_lines_literal_eligible(p->dev) || p->dfa.max_match_len < ugx::SC_REGION - 512;
  // (chunk-wise overlap of copy and scan for page-locked buffers; pageable ones go through the feeder threads of resolve())
  const bool pipelined = stream_route && bounded && !p->never && !s->no_pipeline && n >= 2 * PIPE_CHUNK && is_host_pointer(buf) &&
                         (is_pinned_host_pointer(buf) || s->no_feeder);
  int rc;
  if (pipelined)
  {
    rc = ensure(s->stage, s->stage_cap, n + 16);
    dbuf = s->stage;
    h2d = n;
  }
  else
    rc = resolve(s, buf, n, &dbuf, &h2d);
  if (rc != UGX_OK)
    return rc;
  if (p->never)
  {
    // no position can be a candidate: the result is empty; the pass over the text only counts newlines
    CU(cudaEventRecord(s->ev0, s->stream));
    CU(ugx::launch_count_newlines(dbuf, n, s->totals + 1, s->sm_count, s->stream));
    CU(cudaMemcpyAsync(s->h_totals + 1, s->totals + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev1, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms0 = 0;
    CU(cudaEventElapsedTime(&ms0, s->ev0, s->ev1));
    tt.newlines = s->h_totals[1];
    tt.flags |= UGX_TOT_NEWLINES;
    tt.kernel_ms = ms0;
    tt.launches = 1;
    tt.kernel = UGX_K_NEWLINES;
    s->records_n = 0;
    if (totals)
      *totals = tt;
    return UGX_OK;
  }
  if (mode == 1 && !s->force_generic && !s->no_span && !s->match_lines && !s->two_pass_records &&
      ugx::span_scan_eligible(p->dev))
  {
    bool valid = false;
    CU(cudaEventRecord(s->ev0, s->stream));
    rc = scan_spans(s, p, dbuf, n, want_records, base_offset, base_line, &tt, &valid);
    if (rc != UGX_OK)
      return rc;
    if (valid)
    {
      CU(cudaEventRecord(s->ev1, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      float sms = 0;
      CU(cudaEventElapsedTime(&sms, s->ev0, s->ev1));
      tt.kernel_ms = sms;
      if (want_records)
      {
        if (n_out)
          *n_out = s->records_n;
        if (dev_out)
          *dev_out = s->records;
      }
      if (totals)
        *totals = tt;
      return UGX_OK;
    }
    tt.launches = 0; // the line-at-a-time kernels take over
    tt.flags |= UGX_TOT_SPAN_HANDOVER;
  }
  const uint64_t tile_bytes = ugx::scan_tile_bytes(p->dev);
  const uint64_t ntiles = (n + tile_bytes - 1) / tile_bytes;
  const uint64_t nstrips = (n + ugx::SCAN_STRIP - 1) / ugx::SCAN_STRIP;
  uint64_t cap1 = s->tiles_cap, cap2 = s->tiles_cap;
  rc = ensure(s->tile_matches, cap1, ntiles);
  if (rc == UGX_OK)
    rc = ensure(s->tile_newlines, cap2, ntiles);
  s->tiles_cap = cap1 < cap2 ? cap1 : cap2;
  if (rc == UGX_OK && want_records && s->two_pass_records)
    rc = ensure(s->strip_counts, s->strips_cap, nstrips);
  if (rc != UGX_OK)
    return rc;
  ugx::ScanArgs a;
  a.buf = dbuf;
  a.n = n;
  a.ntiles = ntiles;
  a.tile_matches = s->tile_matches;
  a.tile_newlines = s->tile_newlines;
  a.strip_counts = want_records && s->two_pass_records ? s->strip_counts : nullptr;
  a.out = nullptr;
  a.out_cap = 0;
  a.base_offset = base_offset;
  a.base_line = base_line;
  CU(cudaEventRecord(s->ev0, s->stream));
  if (stream_route)
  {
    rc = ensure(s->region_sum, s->region_cap, ugx::stream_regions(n) + 64);
    if (rc != UGX_OK)
      return rc;
    ugx::StreamArgs sa;
    sa.region_sum = s->region_sum;
    sa.partials = s->partials;
    sa.ticket = s->ticket;
    sa.done = reinterpret_cast<unsigned int*>(s->ticket + 1);
    sa.totals = s->totals;
    sa.stage_table = 0;
    sa.use_h4 = 0;
    sa.use_via = 0;
    sa.no_cover = s->no_cover ? 1u : 0u;
    const uint64_t nreg = ugx::stream_regions(n);
    if (pipelined)
    {
      // host buffer: copy in chunks on the copy stream and scan chunk i while chunk i + 1 is in flight.  A launch
      // trails the copied bytes by one region, so that every read of a scanned position (literal verify, a DFA
      // attempt of bounded length) finds its bytes; the region summaries chain the launches.
      uint64_t copied = 0, scanned = 0;
      tt.launches = 0;
      while (scanned < nreg)
      {
        const uint64_t take = n - copied < PIPE_CHUNK ? n - copied : PIPE_CHUNK;
        if (take > 0)
        {
          CU(cudaMemcpyAsync(s->stage + copied, static_cast<const uint8_t*>(buf) + copied, take, cudaMemcpyHostToDevice,
                             s->copy_stream));
          copied += take;
          CU(cudaEventRecord(s->ev_copy, s->copy_stream));
          CU(cudaStreamWaitEvent(s->stream, s->ev_copy, 0));
        }
        const uint64_t upto = copied == n ? nreg : copied / ugx::SC_REGION - 1;
        if (upto <= scanned)
          continue;
        sa.region_begin = scanned;
        sa.region_end = upto;
        sa.finalize = upto == nreg ? 1u : 0u;
        sa.accumulate = scanned == 0 ? 0u : 1u;
        CU(ugx::launch_count_lines_stream(p->dev, dbuf, copied, sa, s->count_newlines, s->sm_count, s->stream));
        scanned = upto;
        ++tt.launches;
      }
    }
    else
    {
      sa.region_begin = 0;
      sa.region_end = nreg;
      sa.finalize = 1;
      sa.accumulate = 0;
      CU(ugx::launch_count_lines_stream(p->dev, dbuf, n, sa, s->count_newlines, s->sm_count, s->stream));
      tt.launches = 1;
    }
    tt.kernel = ugx::count_lines_literal_eligible(p->dev) ? UGX_K_STREAM_LITERAL : UGX_K_STREAM_DFA;
  }
  else if (mode == 0 && !want_records && !s->force_generic && ugx::count_lines_any_eligible(p->dev))
  {
    CU(ugx::launch_count_lines_any(p->dev, dbuf, n, s->totals, s->sm_count, s->stream));
    tt.launches = 1;
    tt.kernel = UGX_K_TILE_ANY;
  }
  else if (!want_records && !s->force_generic && s->match_lines && ugx::match_lines_eligible(p->dev))
  {
    // counting with position-parallel attempts (match_lines.cu); its tiles are smaller than the line scan's
    const uint64_t tb = ugx::match_lines_tile_bytes(p->dev);
    const uint64_t nt = (n + tb - 1) / tb;
    uint64_t c1 = s->tiles_cap, c2 = s->tiles_cap;
    rc = ensure(s->tile_matches, c1, nt);
    if (rc == UGX_OK)
      rc = ensure(s->tile_newlines, c2, nt);
    s->tiles_cap = c1 < c2 ? c1 : c2;
    if (rc != UGX_OK)
      return rc;
    a.ntiles = nt;
    a.tile_matches = s->tile_matches;
    a.tile_newlines = s->tile_newlines;
    CU(ugx::launch_match_lines(p->dev, a, mode, s->sm_count, s->stream));
    CU(ugx::launch_tile_prefix(s->tile_matches, s->tile_newlines, nt, s->totals, s->stream));
    tt.launches = 2;
    tt.kernel = UGX_K_MATCH_LINES;
  }
  else if (want_records && !s->two_pass_records)
  {
    // single-pass records: staging pass (tiles in completion order) + prefix + reorder into input order.  The staging
    // capacity is a guess (the last result, or one record per 96 bytes); the cursor tells the exact need, so a
    // second staging pass always fits.
    rc = ensure(s->tile_base, s->tile_base_cap, ntiles);
    uint64_t guess = s->records_n + s->records_n / 4 + n / 96 + 1024;
    if (rc == UGX_OK && s->rec_stage_cap < guess)
      rc = ensure(s->rec_stage, s->rec_stage_cap, guess);
    if (rc != UGX_OK)
      return rc;
    unsigned long long* cursor = s->totals + 2;
    uint64_t nrec = 0;
    for (int attempt = 0; attempt < 2; ++attempt)
    {
      CU(ugx::launch_scan_records(p->dev, a, s->tile_base, s->rec_stage, s->rec_stage_cap, cursor, s->sm_count, s->stream));
      CU(ugx::launch_tile_prefix(s->tile_matches, s->tile_newlines, ntiles, s->totals, s->stream));
      CU(cudaMemcpyAsync(s->h_totals, s->totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
      CU(cudaStreamSynchronize(s->stream));
      tt.launches += 2;
      nrec = s->h_totals[0];
      if (nrec <= s->rec_stage_cap)
        break;
      if (attempt == 1)
        return fail(UGX_E_CUDA, "record staging overflowed twice");
      rc = ensure(s->rec_stage, s->rec_stage_cap, nrec);
      if (rc != UGX_OK)
        return rc;
    }
    rc = ensure(s->records, s->records_cap, nrec);
    if (rc != UGX_OK)
      return rc;
    if (nrec > 0)
    {
      CU(ugx::launch_reorder_records(s->rec_stage, s->records, s->tile_matches, s->tile_newlines, s->tile_base, ntiles,
                                     s->totals, base_line, s->sm_count, s->stream));
      tt.launches += 1;
    }
    tt.kernel = UGX_K_RECORDS;
    s->records_n = nrec;
    if (n_out)
      *n_out = nrec;
    if (dev_out)
      *dev_out = s->records;
  }
  else
  {
    CU(ugx::launch_scan_lines(p->dev, a, mode, false, s->sm_count, s->stream));
    CU(ugx::launch_tile_prefix(s->tile_matches, s->tile_newlines, ntiles, s->totals, s->stream));
    tt.launches = 2;
    tt.kernel = UGX_K_LINE_SCAN;
  }
  if (!(want_records && !s->two_pass_records))
    CU(cudaMemcpyAsync(s->h_totals, s->totals, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  if (want_records && s->two_pass_records)
  {
    CU(cudaStreamSynchronize(s->stream));
    const uint64_t nrec = s->h_totals[0];
    rc = ensure(s->records, s->records_cap, nrec);
    if (rc != UGX_OK)
      return rc;
    if (nrec > 0)
    {
      a.out = s->records;
      a.out_cap = nrec;
      CU(ugx::launch_scan_lines(p->dev, a, mode, true, s->sm_count, s->stream));
      tt.launches += 1;
    }
    s->records_n = nrec;
    if (n_out)
      *n_out = nrec;
    if (dev_out)
      *dev_out = s->records;
  }
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  tt.matches = s->h_totals[0];
  tt.newlines = s->h_totals[1];
  if (!stream_route || s->count_newlines)
    tt.flags |= UGX_TOT_NEWLINES;
  tt.kernel_ms = ms;
  if (totals)
    *totals = tt;
  return UGX_OK;
}

} // namespace

extern "C" {

int ugx_count_lines(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t nbytes, ugx_totals* totals)
{
  return scan_common(s, p, buf, nbytes, 0, false, 0, 0, nullptr, nullptr, totals);
}

int ugx_count_matches(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t nbytes, ugx_totals* totals)
{
  return scan_common(s, p, buf, nbytes, 1, false, 0, 0, nullptr, nullptr, totals);
}

int ugx_find_all_device(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t nbytes, uint64_t base_offset,
                        uint64_t base_line, const ugx_match** dev_out, uint64_t* n_out, ugx_totals* totals)
{
  return scan_common(s, p, buf, nbytes, 1, true, base_offset, base_line, dev_out, n_out, totals);
}

int ugx_find_all(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t nbytes, uint64_t base_offset,
                 uint64_t base_line, ugx_match* out, uint64_t cap, uint64_t* n_out, ugx_totals* totals)
{
  const ugx_match* dev = nullptr;
  uint64_t n = 0;
  int rc = scan_common(s, p, buf, nbytes, 1, true, base_offset, base_line, &dev, &n, totals);
  if (rc != UGX_OK)
    return rc;
  if (n_out)
    *n_out = n;
  if (n > cap)
    return fail(UGX_E_OVERFLOW, "record buffer too small");
  if (n > 0)
  {
    if (out == nullptr)
      return fail(UGX_E_INVALID, "null record buffer");
    CU(cudaMemcpyAsync(out, dev, n * sizeof(ugx_match), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
  }
  return UGX_OK;
}

const char* ugx_kernel_name(uint32_t id)
{
  switch (id)
  {
    case UGX_K_STREAM_LITERAL: return "count_lines_literal_kernel";
    case UGX_K_STREAM_DFA: return "count_lines_stream_kernel";
    case UGX_K_TILE_ANY: return "count_lines_any_kernel";
    case UGX_K_LINE_SCAN: return "scan_lines_kernel";
    case UGX_K_RECORDS: return "scan_records_kernel";
    case UGX_K_NEWLINES: return "count_newlines_kernel";
    case UGX_K_MATCH_LINES: return "match_lines_kernel";
    case UGX_K_SPAN: return "span_scan_kernel";
    case UGX_K_BATCH: return "scan_batch_kernel";
    default: return "none";
  }
}

int ugx_scanner_set_option(ugx_scanner* s, const char* name, int value)
{
  if (s == nullptr || name == nullptr)
    return fail(UGX_E_INVALID, "null argument");
  if (strcmp(name, "force_generic") == 0)
  {
    s->force_generic = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "legacy_any") == 0)
  {
    s->legacy_any = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "match_lines") == 0)
  {
    s->match_lines = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "two_pass_records") == 0)
  {
    s->two_pass_records = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "no_pipeline") == 0)
  {
    s->no_pipeline = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "stream_dfa") == 0)
  {
    s->stream_dfa = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "no_feeder") == 0)
  {
    s->no_feeder = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "no_span") == 0)
  {
    s->no_span = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "no_cover") == 0)
  {
    s->no_cover = value != 0;
    return UGX_OK;
  }
  if (strcmp(name, "count_newlines") == 0)
  {
    s->count_newlines = value != 0;
    return UGX_OK;
  }
  return fail(UGX_E_INVALID, std::string("unknown scanner option ") + name);
}

int ugx_scanner_fetch(ugx_scanner* s, ugx_match* out, uint64_t first, uint64_t count)
{
  if (s == nullptr || (out == nullptr && count != 0))
    return fail(UGX_E_INVALID, "null argument");
  if (first > s->records_n || count > s->records_n - first)
    return fail(UGX_E_INVALID, "record range outside the last result");
  if (count == 0)
    return UGX_OK;
  CU(cudaSetDevice(s->device));
  CU(cudaMemcpyAsync(out, s->records + first, count * sizeof(ugx_match), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  return UGX_OK;
}

int ugx_count_batch(ugx_scanner* s, const ugx_pattern* p, const void* buf, uint64_t nbytes, const uint64_t* begins,
                    const uint64_t* lens, uint64_t nfiles, int mode, uint64_t* counts, ugx_totals* totals)
{
  if (s == nullptr || p == nullptr || (buf == nullptr && nbytes != 0) || (nfiles != 0 && (begins == nullptr || lens == nullptr || counts == nullptr)) ||
      (mode != UGX_MODE_LINES && mode != UGX_MODE_MATCHES))
    return fail(UGX_E_INVALID, "ugx_count_batch: bad argument");
  if (s->device != p->device)
    return fail(UGX_E_INVALID, "pattern and scanner live on different devices");
  ugx_totals tt;
  memset(&tt, 0, sizeof(tt));
  if (totals)
    *totals = tt;
  if (nfiles == 0)
    return UGX_OK;
  try
  {
    // the tile table: tile g of the launch is tile tile_index[g] of file tile_file[g]
    std::vector<uint32_t> tf, ti;
    for (uint64_t f = 0; f < nfiles; ++f)
    {
      if ((begins[f] & 15) != 0 || begins[f] > nbytes || lens[f] > nbytes - begins[f])
        return fail(UGX_E_INVALID, "ugx_count_batch: file " + std::to_string(f) + " is not a 16-byte aligned range of the buffer");
      if (nfiles > 0xffffffffull || lens[f] / ugx::SCAN_TILE > 0xfffffffeull)
        return fail(UGX_E_INVALID, "ugx_count_batch: too many files / a file too large for one batch");
      const uint64_t nt = (lens[f] + ugx::SCAN_TILE - 1) / ugx::SCAN_TILE;
      for (uint64_t j = 0; j < nt; ++j)
      {
        tf.push_back(static_cast<uint32_t>(f));
        ti.push_back(static_cast<uint32_t>(j));
      }
    }
    memset(counts, 0, nfiles * sizeof(uint64_t));
    if (tf.empty() || p->never)
      return UGX_OK; // only empty files, or a prefilter that admits no candidate (config 3): nothing matches
    CU(cudaSetDevice(s->device));
    const uint8_t* dbuf = nullptr;
    uint64_t h2d = 0;
    int rc = resolve(s, buf, nbytes, &dbuf, &h2d);
    if (rc != UGX_OK)
      return rc;
    // one scratch block: begins | lens | counts | tile_file | tile_index
    const uint64_t words = 3 * nfiles + (tf.size() + 1) / 2 * 2;
    rc = ensure(s->batch, s->batch_cap, words + 4);
    if (rc != UGX_OK)
      return rc;
    uint64_t* d_begins = s->batch;
    uint64_t* d_lens = d_begins + nfiles;
    unsigned long long* d_counts = reinterpret_cast<unsigned long long*>(d_lens + nfiles);
    uint32_t* d_tf = reinterpret_cast<uint32_t*>(d_counts + nfiles);
    uint32_t* d_ti = d_tf + tf.size();
    CU(cudaMemcpyAsync(d_begins, begins, nfiles * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(d_lens, lens, nfiles * 8, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemsetAsync(d_counts, 0, nfiles * 8, s->stream));
    CU(cudaMemcpyAsync(d_tf, tf.data(), tf.size() * 4, cudaMemcpyHostToDevice, s->stream));
    CU(cudaMemcpyAsync(d_ti, ti.data(), ti.size() * 4, cudaMemcpyHostToDevice, s->stream));
    ugx::BatchArgs a;
    a.begins = d_begins;
    a.lens = d_lens;
    a.tile_file = d_tf;
    a.tile_index = d_ti;
    a.ntiles = tf.size();
    a.counts = d_counts;
    a.stage_table = 0;
    CU(cudaEventRecord(s->ev0, s->stream));
    CU(ugx::launch_scan_batch(p->dev, dbuf, a, mode, s->sm_count, s->stream));
    CU(cudaEventRecord(s->ev1, s->stream));
    CU(cudaMemcpyAsync(counts, d_counts, nfiles * 8, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    for (uint64_t f = 0; f < nfiles; ++f)
      tt.matches += counts[f];
    tt.kernel_ms = ms;
    tt.launches = 1;
    tt.kernel = UGX_K_BATCH;
    if (totals)
      *totals = tt;
    return UGX_OK;
  }
  catch (const std::bad_alloc&)
  {
    return fail(UGX_E_NOMEM, "out of host memory");
  }
}

int ugx_check_text(ugx_scanner* s, const void* buf, uint64_t nbytes, ugx_text_info* out)
{
  if (s == nullptr || out == nullptr || (buf == nullptr && nbytes != 0))
    return fail(UGX_E_INVALID, "null argument");
  memset(out, 0, sizeof(*out));
  out->is_utf8 = 1;
  if (nbytes == 0)
    return UGX_OK;
  CU(cudaSetDevice(s->device));
  const uint8_t* dbuf = nullptr;
  uint64_t h2d = 0;
  const int rc = resolve(s, buf, nbytes, &dbuf, &h2d);
  if (rc != UGX_OK)
    return rc;
  unsigned int* dflags = reinterpret_cast<unsigned int*>(s->totals + 7);
  CU(cudaMemsetAsync(dflags, 0, sizeof(unsigned long long), s->stream));
  CU(cudaEventRecord(s->ev0, s->stream));
  CU(ugx::launch_utf8_check(dbuf, nbytes, dflags, s->sm_count, s->stream));
  CU(cudaMemcpyAsync(s->h_totals + 7, s->totals + 7, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
  CU(cudaEventRecord(s->ev1, s->stream));
  CU(cudaStreamSynchronize(s->stream));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
  const unsigned int f = static_cast<unsigned int>(s->h_totals[7]);
  out->is_utf8 = (f & 1u) == 0;
  out->has_nul = (f & 2u) != 0;
  out->kernel_ms = ms;
  out->launches = 1;
  return UGX_OK;
}

int ugx_count_newlines(ugx_scanner* s, const void* buf, uint64_t nbytes, ugx_totals* totals)
{
  if (s == nullptr || (buf == nullptr && nbytes != 0))
    return fail(UGX_E_INVALID, "null argument");
  ugx_totals tt;
  memset(&tt, 0, sizeof(tt));
  if (nbytes != 0)
  {
    CU(cudaSetDevice(s->device));
    const uint8_t* dbuf = nullptr;
    uint64_t h2d = 0;
    const int rc = resolve(s, buf, nbytes, &dbuf, &h2d);
    if (rc != UGX_OK)
      return rc;
    CU(cudaEventRecord(s->ev0, s->stream));
    CU(ugx::launch_count_newlines(dbuf, nbytes, s->totals + 1, s->sm_count, s->stream));
    CU(cudaMemcpyAsync(s->h_totals + 1, s->totals + 1, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaEventRecord(s->ev1, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    tt.newlines = s->h_totals[1];
    tt.kernel_ms = ms;
    tt.launches = 1;
    tt.kernel = UGX_K_NEWLINES;
  }
  tt.flags |= UGX_TOT_NEWLINES;
  if (totals)
    *totals = tt;
  return UGX_OK;
}

} // extern "C"
