// tests/csrc/simt_emu.cpp — TEST-ONLY warp emulator, see simt_emu.h
#include "simt_emu.h"

namespace emu {

thread_local Warp* g_warp = nullptr;
static thread_local uint64_t g_total_collectives = 0;
static uint64_t g_site_count[4096];     // EMU_SITES=1: how often each call site (source line) ran

// save the callee-saved registers on the current stack, switch stacks, restore (System V x86-64)
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

static const size_t kStack = 512 * 1024;

static void trampoline() {
    Warp* w = g_warp;
    const int lane = w->cur;
    w->entry(lane, w->user);
    w->done[lane] = true;
    w->op[lane] = OP_NONE;
    for (;;) emu_switch(&w->lane_sp[lane], w->sched_sp);    // never resumed
}

static const char* op_name(int op) {
    static const char* n[] = {"none", "ballot", "any", "all", "shfl", "shfl_xor", "match_any", "reduce_xor",
                              "reduce_or", "reduce_add", "reduce_min", "reduce_max", "syncwarp"};
    return n[op];
}

static void resolve(Warp* w) {
    // uniform control flow: every lane sits at the same site with the same operation
    for (int l = 0; l < 32; l++) {
        if (w->done[l] || w->site[l] != w->site[0] || w->op[l] != w->op[0]) {
            fprintf(stderr, "simt_emu: divergent collective: lane 0 at line %d (%s), lane %d %s line %d (%s)\n",
                    w->site[0], op_name(w->op[0]), l, w->done[l] ? "EXITED, last at" : "at", w->site[l],
                    op_name(w->op[l]));
            abort();
        }
    }
    for (int l = 0; l < 32; l++) {
        const unsigned m = w->mask[l];
        if (!((m >> l) & 1u)) {
            fprintf(stderr, "simt_emu: lane %d not in its own member mask %08x at line %d\n", l, m, w->site[l]);
            abort();
        }
        uint64_t r = 0;
        switch (w->op[l]) {
            case OP_BALLOT:
                for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && w->val[k]) r |= 1ull << k;
                break;
            case OP_ANY:
                for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && w->val[k]) r = 1;
                break;
            case OP_ALL:
                r = 1;
                for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && !w->val[k]) r = 0;
                break;
            case OP_SHFL: {
                const int src = w->arg[l] & 31;
                r = w->val[src];       // reading a lane outside the mask is undefined on the GPU
                if (!((m >> src) & 1u)) {
                    fprintf(stderr, "simt_emu: shfl from lane %d outside mask %08x at line %d\n", src, m, w->site[l]);
                    abort();
                }
                break;
            }
            case OP_SHFL_XOR: {
                const int src = (l ^ w->arg[l]) & 31;
                r = ((m >> src) & 1u) ? w->val[src] : w->val[l];
                break;
            }
            case OP_MATCH:
                for (int k = 0; k < 32; k++) if (((m >> k) & 1u) && w->val[k] == w->val[l]) r |= 1ull << k;
                break;
            case OP_RED_XOR:
                for (int k = 0; k < 32; k++) if ((m >> k) & 1u) r ^= w->val[k];
                break;
            case OP_RED_OR:
                for (int k = 0; k < 32; k++) if ((m >> k) & 1u) r |= w->val[k];
                break;
            case OP_RED_ADD:
                for (int k = 0; k < 32; k++) if ((m >> k) & 1u) r += w->val[k];
                r &= 0xffffffffull;
                break;
            case OP_RED_MIN:
                r = 0xffffffffull;
                for (int k = 0; k < 32; k++) if ((m >> k) & 1u) r = std::min(r, w->val[k]);
                break;
            case OP_RED_MAX:
                for (int k = 0; k < 32; k++) if ((m >> k) & 1u) r = std::max(r, w->val[k]);
                break;
            case OP_SYNC:
                break;
            default:
                fprintf(stderr, "simt_emu: bad op\n");
                abort();
        }
        w->res[l] = r;
    }
    // the lanes named by a mask must agree on it (each group of a partitioned warp passes its own)
    for (int l = 0; l < 32; l++)
        for (int k = 0; k < 32; k++)
            if (((w->mask[l] >> k) & 1u) && w->mask[k] != w->mask[l]) {
                fprintf(stderr, "simt_emu: lanes %d and %d disagree on the member mask at line %d\n", l, k, w->site[l]);
                abort();
            }
    w->n_collectives++;
    g_site_count[w->site[0] & 4095]++;
    w->trace[w->n_collectives & 63] = w->site[0];
    if (w->limit && w->n_collectives > w->limit) {
        fprintf(stderr, "simt_emu: more than %llu collectives in one warp (EMU_MAX_COLLECTIVES); last sites:", (unsigned long long)w->limit);
        for (int i = 63; i >= 0; i--) fprintf(stderr, " %d", w->trace[(w->n_collectives - i) & 63]);
        fprintf(stderr, "\n");
        abort();
    }
}

uint64_t collective(int op, int site, unsigned mask, uint64_t val, int arg) {
    Warp* w = g_warp;
    const int l = w->cur;
    w->op[l] = op; w->site[l] = site; w->mask[l] = mask; w->val[l] = val; w->arg[l] = arg;
    emu_switch(&w->lane_sp[l], w->sched_sp);
    return w->res[l];
}

uint64_t collectives_run() { return g_total_collectives; }

void run_warp(void (*entry)(int lane, void* user), void* user) {
    Warp* w = new Warp;
    memset(w, 0, sizeof(*w));
    w->entry = entry;
    w->user = user;
    const char* sh = getenv("EMU_SHUFFLE");
    w->shuffle = sh && atoi(sh) != 0;
    w->rng = sh ? (uint32_t)atoi(sh) * 2654435761u + 12345u : 1u;
    const char* lim = getenv("EMU_MAX_COLLECTIVES");
    w->limit = lim ? strtoull(lim, nullptr, 10) : 0;
    for (int l = 0; l < 32; l++) {
        w->stack[l] = (char*)aligned_alloc(64, kStack);
        uintptr_t top = ((uintptr_t)(w->stack[l] + kStack)) & ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;                    // return address of the trampoline (never used)
        *--sp = (void*)&trampoline;         // popped by the `ret` of emu_switch
        for (int k = 0; k < 6; k++) *--sp = nullptr;
        w->lane_sp[l] = sp;
    }
    Warp* saved = g_warp;
    g_warp = w;
    int order[32];
    for (int l = 0; l < 32; l++) order[l] = l;
    for (;;) {
        if (w->shuffle) {
            for (int i = 31; i > 0; i--) {
                w->rng = w->rng * 1664525u + 1013904223u;
                const int j = (int)((w->rng >> 8) % (uint32_t)(i + 1));
                std::swap(order[i], order[j]);
            }
        }
        int live = 0;
        for (int i = 0; i < 32; i++) {
            const int l = order[i];
            if (w->done[l]) continue;
            w->cur = l;
            emu_switch(&w->sched_sp, w->lane_sp[l]);
            if (!w->done[l]) live++;
        }
        if (live == 0) break;
        resolve(w);
    }
    g_total_collectives += w->n_collectives;
    if (getenv("EMU_SITES")) {
        for (int i = 0; i < 4096; i++)
            if (g_site_count[i]) fprintf(stderr, "site line %d: %llu\n", i, (unsigned long long)g_site_count[i]);
        memset(g_site_count, 0, sizeof(g_site_count));
    }
    g_warp = saved;
    for (int l = 0; l < 32; l++) free(w->stack[l]);
    delete w;
}

}  // namespace emu
