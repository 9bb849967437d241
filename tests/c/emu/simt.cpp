// Scheduler of the cooperative SIMT interpreter (see simt.h).  Test infrastructure.
#include "simt.h"

uint3_emu threadIdx, blockIdx;
dim3 blockDim, gridDim;

asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

Cta g_cta;
unsigned g_cur = 0;
void* g_sched_sp = nullptr;

namespace {
constexpr size_t kStackBytes = 256 * 1024;
std::vector<char*> g_stacks;

void lane_entry() {
    (*g_cta.body)();
    g_cta.lanes[g_cur].state = DONE;
    void* dead = nullptr;
    emu_switch(&dead, g_sched_sp);
    abort();   // a finished fiber is never resumed
}

void prepare_lane(unsigned tid) {
    while (g_stacks.size() <= tid) {
        void* p = nullptr;
        if (posix_memalign(&p, 64, kStackBytes)) abort();
        g_stacks.push_back((char*)p);
    }
    uintptr_t top = ((uintptr_t)g_stacks[tid] + kStackBytes) & ~(uintptr_t)15;
    void** p = (void**)top;
    *(--p) = nullptr;                      // fake return address of lane_entry (never used)
    *(--p) = (void*)&lane_entry;           // `ret` of emu_switch jumps here; rsp is then 8 mod 16 as after a call
    for (int i = 0; i < 6; i++) *(--p) = nullptr;   // rbp rbx r12 r13 r14 r15
    g_cta.lanes[tid].sp = (void*)p;
    g_cta.lanes[tid].state = READY;
    g_cta.lanes[tid].seq = 0;
}

void resume(unsigned tid) {
    set_thread(tid);
    emu_switch(&g_sched_sp, g_cta.lanes[tid].sp);
}

void run_cta() {
    const unsigned n = g_cta.nthreads, nwarps = (n + 31u) / 32u;
    while (true) {
        bool any_cta_wait = false, all_done = true;
        for (unsigned w = 0; w < nwarps; w++) {
            const unsigned lo = w * 32u, hi = lo + 32u < n ? lo + 32u : n;
            while (true) {
                for (unsigned t = lo; t < hi; t++)
                    if (g_cta.lanes[t].state == READY) resume(t);
                unsigned n_warp = 0, n_ctaw = 0, n_live = 0;
                uint32_t seq = 0;
                bool seq_mismatch = false;
                for (unsigned t = lo; t < hi; t++) {
                    const Lane& L = g_cta.lanes[t];
                    if (L.state == DONE) continue;
                    n_live++;
                    if (L.state == WAIT_WARP) { if (n_warp && L.seq != seq) seq_mismatch = true; seq = L.seq; n_warp++; }
                    if (L.state == WAIT_CTA) n_ctaw++;
                }
                if (n_warp && n_ctaw) { fprintf(stderr, "[emu] warp %u: lanes wait at a warp collective AND at __syncthreads\n", w); abort(); }
                if (seq_mismatch) { fprintf(stderr, "[emu] warp %u: lanes wait at different warp collectives (divergent *_sync)\n", w); abort(); }
                if (n_warp) { for (unsigned t = lo; t < hi; t++) if (g_cta.lanes[t].state == WAIT_WARP) g_cta.lanes[t].state = READY; continue; }
                if (n_ctaw) any_cta_wait = true;
                if (n_live) all_done = false;
                break;
            }
        }
        if (all_done) return;
        if (!any_cta_wait) { fprintf(stderr, "[emu] scheduler stalled\n"); abort(); }
        for (unsigned t = 0; t < n; t++)
            if (g_cta.lanes[t].state == WAIT_CTA) g_cta.lanes[t].state = READY;
    }
}
}  // namespace

void launch(dim3 grid, dim3 block, size_t, const std::function<void()>& body) {
    gridDim = grid;
    blockDim = block;
    g_cta.nthreads = block.x * block.y * block.z;
    g_cta.lanes.assign(g_cta.nthreads, Lane());
    g_cta.xbuf.assign((size_t)((g_cta.nthreads + 31u) / 32u) * 2u * 32u, 0);
    g_cta.body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
                for (unsigned t = 0; t < g_cta.nthreads; t++) prepare_lane(t);
                run_cta();
            }
}

}  // namespace emu
