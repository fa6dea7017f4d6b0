// trace.cuh — optional device-side timeline (compiled only with -DFLB_TRACE into a separate debug library by
// tools/trace_build.sh; the product library contains none of it).  Each instrumented kernel records the earliest block
// start and the latest block end on the global timer into a slot (kernel id x pass), so the real critical path of a
// graph-launched scan — including launch gaps and side-stream overlap — can be read back without a profiler.
#pragma once
#ifdef FLB_TRACE
namespace flb {
constexpr int TRACE_SLOTS = 128;
constexpr int TRACE_PHASES = 96;
struct TraceRec { unsigned long long t0, t1; };
__device__ TraceRec g_trace[TRACE_SLOTS];
__device__ long long g_phase_clk[TRACE_PHASES];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_mark(int slot, bool begin) {
  if (threadIdx.x == 0) {
    const unsigned long long t = gtimer();
    if (begin) atomicMin(&g_trace[slot].t0, t);
    else atomicMax(&g_trace[slot].t1, t);
  }
}
constexpr int TRACE_DBG = 64;
__device__ unsigned long long g_dbg[TRACE_DBG];
__device__ __forceinline__ void dbg_add(int idx, unsigned long long v) { atomicAdd(&g_dbg[idx], v); }
__device__ __forceinline__ void dbg_max(int idx, unsigned long long v) { atomicMax(&g_dbg[idx], v); }
__device__ __forceinline__ void trace_phase(int idx) {
  if (threadIdx.x == 0 && idx < TRACE_PHASES) g_phase_clk[idx] = clock64();
}
}  // namespace flb
#define FLB_TRACE_BEGIN(slot) flb::trace_mark((slot), true)
#define FLB_TRACE_END(slot) flb::trace_mark((slot), false)
#define FLB_TRACE_PHASE(idx) flb::trace_phase(idx)
#define FLB_DBG_ADD(idx, v) flb::dbg_add((idx), (unsigned long long)(v))
#define FLB_DBG_MAX(idx, v) flb::dbg_max((idx), (unsigned long long)(v))
#define FLB_DBG_CLOCK(var) const long long var = clock64()
#else
#define FLB_TRACE_BEGIN(slot)
#define FLB_TRACE_END(slot)
#define FLB_TRACE_PHASE(idx)
#define FLB_DBG_ADD(idx, v)
#define FLB_DBG_MAX(idx, v)
#define FLB_DBG_CLOCK(var)
#endif
