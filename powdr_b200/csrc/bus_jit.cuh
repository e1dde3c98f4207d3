// Stage 0, periphery histograms, generated per AIR: the drop-in `_apc_apply_bus` keeps the reference's shape --
// one thread per row walks all interactions serially through a bytecode interpreter with a local-memory stack
// (/root/reference/openvm/cuda/src/apc_apply_bus.cu:23-113).  An autoprecompile's interactions are fixed at key-generation time, so
// pb_bus_compile turns them into straight-line CUDA C with the same SSA emitter as the LogUp kernels (constants folded, column loads
// hoisted per group) and pb_bus_apply launches a (row tile) x (interaction group) grid.  Semantics = cpu/periphery.rs:179-236:
//   variable range checker  [value, max_bits]            -> bin (1 << max_bits) + value - 1
//   tuple range checker     [v0, v1]                     -> bin v0 * sz1 + v1
//   bitwise lookup          [x, y, x^y, selector]        -> bin (x << 8 | y) (+ 2^16 when selector = 1)   (count layout [range | xor])
// other buses (execution bridge, memory, pc lookup) are skipped; multiplicities are added with one warp-aggregated atomic per bin.
#pragma once
#include "logup_jit.cuh"

namespace busjit {

struct Kernel {
    std::vector<CUmodule> mods;
    std::vector<CUfunction> fns;
    size_t n_periphery = 0;
};

static const char* BUS_PRELUDE = R"(
__device__ __forceinline__ u32 canon(u32 m) { return mul(m, 1u); }          // Montgomery -> canonical
__device__ __forceinline__ void hist_add(u32* hist, u32 idx, u32 mult) {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, idx);
    const int leader = __ffs(peers) - 1;
    u32 total = 0;
    for (unsigned rem = peers; rem;) { int l = __ffs(rem) - 1; rem &= rem - 1; total += __shfl_sync(peers, mult, l); }
    if ((int)(threadIdx.x & 31) == leader) atomicAdd(hist + idx, total);
}
)";

constexpr size_t GROUP = 16;     // periphery interactions per generated function / blockIdx.y slice

inline std::string generate(const std::vector<uint32_t>& code, const std::vector<air::Span>& spans, const std::vector<uint32_t>& pool,
                            const std::vector<logup::Interaction>& ints, size_t g_begin, size_t g_end, uint32_t var_bus, uint32_t t2_bus, uint32_t bw_bus) {
    std::string src = airjit::PRELUDE;
    src += BUS_PRELUDE;
    char buf[512];
    size_t n_fn = 0;
    std::string calls;
    for (size_t g0 = g_begin; g0 < g_end; g0 += GROUP, n_fn++) {
        const size_t g1 = std::min(g_end, g0 + GROUP);
        std::string body, loads;
        std::unordered_map<uint32_t, std::string> cols;
        size_t vid = 0;
        for (size_t i = g0; i < g1; i++) {
            const logup::Interaction& it = ints[i];
            auto val = [&](uint32_t k) {
                logup::Val v = logup::emit_expr(body, code, spans[it.span0 + k], pool, vid, cols, loads);
                if (v.lit) { snprintf(buf, sizeof buf, "%uu", bb::from_monty(v.mont)); return std::string(buf); }
                return "canon(" + v.name + ")";
            };
            const std::string m = val(0);
            if (it.bus_id == var_bus && it.num_args >= 2) {
                const std::string a = val(1), b = val(2);
                snprintf(buf, sizeof buf, " { const u32 m = %s; if (m) { const u32 idx = (1u << (%s & 31u)) + %s - 1u; if (idx < var_bins) hist_add(var_hist, idx, m); } }\n", m.c_str(),
                         b.c_str(), a.c_str());
            } else if (it.bus_id == t2_bus && it.num_args >= 2) {
                const std::string a = val(1), b = val(2);
                snprintf(buf, sizeof buf, " { const u32 m = %s; if (m) { const u32 idx = %s * sz1 + %s; if (idx < sz0 * sz1) hist_add(t2_hist, idx, m); } }\n", m.c_str(), a.c_str(), b.c_str());
            } else if (it.bus_id == bw_bus && it.num_args >= 4) {
                const std::string x = val(1), y = val(2), sel = val(4);
                snprintf(buf, sizeof buf, " { const u32 m = %s; if (m) { const u32 sel = %s; const u32 idx = ((%s << 8) | %s) + (sel == 1u ? 65536u : 0u); if (sel <= 1u && idx < 131072u) hist_add(bw_hist, idx, m); } }\n",
                         m.c_str(), sel.c_str(), x.c_str(), y.c_str());
            } else continue;
            body += buf;
        }
        snprintf(buf, sizeof buf, "__device__ __noinline__ void bg%zu(const u32* __restrict__ b, u64 h, u32* var_hist, u32 var_bins, u32* t2_hist, u32 sz0, u32 sz1, u32* bw_hist) {\n", n_fn);
        src += buf;
        src += loads + body + "}\n";
        snprintf(buf, sizeof buf, "    if (blockIdx.y == %zu) bg%zu(b, h, var_hist, var_bins, t2_hist, sz0, sz1, bw_hist);\n", n_fn, n_fn);
        calls += buf;
    }
    src += R"(
extern "C" __global__ void __launch_bounds__(128) pb_bus(const u32* __restrict__ trace, u64 h, int num_calls, u32* var_hist, u32 var_bins, u32* t2_hist, u32 sz0,
                                                         u32 sz1, u32* bw_hist) {
    const u64 r = (u64)blockIdx.x * 128ull + threadIdx.x;
    if (r >= (u64)num_calls) return;
    const u32* b = trace + r;
)";
    src += calls + "}\n";
    return src;
}

// periphery interactions only (the others never touch a histogram); modules of MOD_GROUPS functions, one grid.y slice per function
constexpr size_t MOD_GROUPS = 8;

inline int build(const std::vector<uint32_t>& code, const std::vector<air::Span>& spans, const std::vector<uint32_t>& pool,
                 const std::vector<logup::Interaction>& all, uint32_t var_bus, uint32_t t2_bus, uint32_t bw_bus, Kernel* out, std::vector<unsigned>* grid_y) {
    airjit::Api& a = airjit::api();
    if (!a.nvrtc_ok || (out && !a.ok)) return 3;                 // out == nullptr: compile only (host-side check, no device needed)
    std::vector<logup::Interaction> ints;
    for (const auto& it : all)
        if (it.bus_id == var_bus || it.bus_id == t2_bus || it.bus_id == bw_bus) ints.push_back(it);
    if (out) out->n_periphery = ints.size();
    if (ints.empty()) return 0;
    const size_t per_mod = GROUP * MOD_GROUPS, n_mods = (ints.size() + per_mod - 1) / per_mod;
    std::vector<std::vector<char>> cubins(n_mods);
    std::vector<int> rcs(n_mods, 0);
    std::atomic<size_t> next{0};
    auto worker = [&]() {
        for (size_t k = next.fetch_add(1); k < n_mods; k = next.fetch_add(1)) {
            const std::string src = generate(code, spans, pool, ints, k * per_mod, std::min(ints.size(), (k + 1) * per_mod), var_bus, t2_bus, bw_bus);
            rcs[k] = airjit::compile_chunk(src, cubins[k]);
        }
    };
    size_t n_threads = std::min<size_t>(n_mods, std::max(1u, std::thread::hardware_concurrency()));
    std::vector<std::thread> ths;
    for (size_t t = 1; t < n_threads; t++) ths.emplace_back(worker);
    worker();
    for (auto& t : ths) t.join();
    for (size_t k = 0; k < n_mods; k++) {
        if (rcs[k]) return rcs[k];
        if (!out) continue;
        CUmodule mod = nullptr;
        CUfunction fn = nullptr;
        if (a.ModuleLoadData(&mod, cubins[k].data()) != CUDA_SUCCESS || a.ModuleGetFunction(&fn, mod, "pb_bus") != CUDA_SUCCESS) return 6;
        out->mods.push_back(mod);
        out->fns.push_back(fn);
        const size_t n_here = std::min(ints.size(), (k + 1) * per_mod) - k * per_mod;
        grid_y->push_back((unsigned)((n_here + GROUP - 1) / GROUP));
    }
    return 0;
}

inline void destroy(Kernel& k) {
    for (CUmodule m : k.mods) airjit::api().ModuleUnload(m);
    k.mods.clear(); k.fns.clear();
}

}  // namespace busjit
