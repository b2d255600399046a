// Shape-specialised micro-subtree kernels.
//
// The interpreter in tiles.cu (k_micro) walks a template with dynamic indices, which forces the per-thread scratch
// into shared memory and costs ~2 800 instructions per subtree.  All micro subtrees of one class have the SAME
// shape, so at upload time this file writes the sweep of that shape out as straight-line CUDA C++ -- every index a
// literal, every intermediate a named double -- and compiles it for sm_100a with NVRTC (--fmad=false, like the rest
// of the library).  The result keeps the whole subtree in registers, issues all of its loads up front, and runs
// ~6x fewer instructions.  The arithmetic, operand for operand, is that of k_micro / cfr.cu / the reference's
// cfr_recursive (rlpoker/cfr/cfr.py:158-255).  RLP_NO_JIT=1 selects the interpreter kernels instead; a missing NVRTC or
// a failed compile is a hard error (no silent fallback).
#include <dlfcn.h>
#include <nvrtc.h>

#include <sstream>
#include <string>

#include "cfr.cuh"

using rlp::fail;

namespace rlp { thread_local int g_jit_compiled = 0, g_jit_failed = 0; }

namespace {

struct Nvrtc {
    void *handle = nullptr;
    bool tried = false;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    bool load() {
        if (tried) return handle != nullptr;
        tried = true;
        for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
            handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
        }
        if (!handle) return false;
#define RLP_SYM(field, sym) field = (decltype(field))dlsym(handle, #sym); if (!field) { handle = nullptr; return false; }
        RLP_SYM(create, nvrtcCreateProgram)
        RLP_SYM(compile, nvrtcCompileProgram)
        RLP_SYM(cubin_size, nvrtcGetCUBINSize)
        RLP_SYM(cubin, nvrtcGetCUBIN)
        RLP_SYM(log_size, nvrtcGetProgramLogSize)
        RLP_SYM(log, nvrtcGetProgramLog)
        RLP_SYM(destroy, nvrtcDestroyProgram)
#undef RLP_SYM
        return true;
    }
};
Nvrtc g_nvrtc;

// Straight-line sweep of one shape.  Names: cb/sr/so{r} = column base / slots of decision node r; s{j} = strategy
// probability of the edge into local node j; x{j} (y{j}) = player-1 (player-2) value of node j; q1_{r}, q2_{r} reach.
// Resident CTAs per SM requested for the micro kernel (caps registers: 8 -> 64 regs, one wave for Leduc-13).
int micro_min_blocks() {
    const char *e = getenv("RLP_MICRO_MINBLOCKS");
    int v = e ? atoi(e) : 5;
    return v < 1 ? 1 : (v > 16 ? 16 : v);
}

bool want_timeline() {
    const char *e = getenv("RLP_TIMELINE");
    return e && e[0] == '1';
}

std::string generate(const MicroTmpl &tm, bool zs) {
    std::ostringstream o;
    const bool tl = want_timeline();
    o.precision(17);
    o << "struct IterState { long long t; int linear; int pad; unsigned long long touched; unsigned long long* tl; unsigned int tl_count, tl_cap; };\n"
         "__device__ __forceinline__ unsigned long long rlp_now() { unsigned long long t; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(t)); return t; }\n"
         "__device__ __forceinline__ int tl_begin(const IterState* it, int kind) {\n"
         "  if (!it->tl || blockIdx.x != 0 || threadIdx.x != 0) return -1;\n"
         "  IterState* m = const_cast<IterState*>(it); unsigned int slot = atomicAdd(&m->tl_count, 1u);\n"
         "  if (slot >= m->tl_cap) return -1; m->tl[3 * slot] = (unsigned long long)kind; m->tl[3 * slot + 1] = rlp_now(); return (int)slot; }\n"
         "__device__ __forceinline__ void tl_end(const IterState* it, int slot) { if (slot >= 0) const_cast<IterState*>(it)->tl[3 * slot + 2] = rlp_now(); }\n"
         "extern \"C\" __global__ void __launch_bounds__(128, " << micro_min_blocks() << ") micro_k(\n"
         "    const int* __restrict__ colbase, const int* __restrict__ slot_r, const int* __restrict__ slot_o,\n"
         "    const int* __restrict__ pidx, const int* __restrict__ ridx,\n"
         "    const double* __restrict__ pay1, const double* __restrict__ pay2, const double* __restrict__ pc,\n"
         "    const double* __restrict__ sigma, const double* __restrict__ portal_p1, const double* __restrict__ portal_p2,\n"
         "    double* __restrict__ portal_v1, double* __restrict__ portal_v2, double* __restrict__ contrib_r,\n"
         "    double* __restrict__ contrib_o, const IterState* __restrict__ iter, long long P,\n"
         "    const int* __restrict__ inst, int count, long long ND, long long NI) {\n"
         "  const int i = blockIdx.x * blockDim.x + threadIdx.x;\n"
         "  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");\n"
         "  if (i >= count) { asm volatile(\"griddepcontrol.wait;\" ::: \"memory\"); return; }\n"
         "  const long long p = inst ? (long long)inst[i] : (long long)i;\n";
    const int n = tm.n, n_nt = tm.n_nt;
    for (int r = 0; r < n_nt; ++r)
        o << "  const int sk" << r << " = colbase[" << r << " * P + p], so" << r << " = slot_o[" << r << " * P + p];\n";
    o << "  const double pcv = pc[p];\n  const int pi = pidx[p], ri = ridx[p];\n";
    for (int j = 0; j < n; ++j)
        if (tm.rank[j] & 0x80) {
            int tr = tm.rank[j] & 0x7F;
            o << "  const double x" << j << " = pay1[" << tr << " * P + p];\n";
            if (!zs) o << "  const double y" << j << " = pay2[" << tr << " * P + p];\n";
        }
    // everything above is static topology; below depends on the previous kernels (strategy, portal reach, counter)
    o << "  asm volatile(\"griddepcontrol.wait;\" ::: \"memory\");\n"
      << (tl ? "  const int tls = tl_begin(iter, 1);\n" : "")
      << "  const double w = iter->linear ? (double)iter->t : 1.0;\n"
         "  const double q1_0 = portal_p1[ri], q2_0 = portal_p2[ri];\n";
    for (int j = 1; j < n; ++j)
        o << "  const double s" << j << " = sigma[" << (int)tm.par_slot[j] << " * NI + sk" << (int)tm.par_rank[j] << "];\n";
    // forward
    for (int r = 0; r < n_nt; ++r) {
        int j = tm.nt_node[r], pl = tm.player[j];
        for (int k = 0; k < tm.nchild[j]; ++k) {
            int c = tm.first_child[j] + k;
            if (tm.rank[c] & 0x80) continue;
            int rc = tm.rank[c];
            o << "  const double q1_" << rc << " = " << (pl == 1 ? "s" + std::to_string(c) + " * " : "") << "q1_" << r << ";\n";
            o << "  const double q2_" << rc << " = " << (pl == 2 ? "s" + std::to_string(c) + " * " : "") << "q2_" << r << ";\n";
        }
    }
    // backward
    for (int r = n_nt - 1; r >= 0; --r) {
        int j = tm.nt_node[r], pl = tm.player[j], fc = tm.first_child[j], nc = tm.nchild[j];
        o << "  double a" << j << " = 0.0;\n";
        for (int k = 0; k < nc; ++k) o << "  a" << j << " = a" << j << " + s" << (fc + k) << " * x" << (fc + k) << ";\n";
        o << "  const double x" << j << " = a" << j << ";\n";
        if (!zs) {
            o << "  double b" << j << " = 0.0;\n";
            for (int k = 0; k < nc; ++k) o << "  b" << j << " = b" << j << " + s" << (fc + k) << " * y" << (fc + k) << ";\n";
            o << "  const double y" << j << " = b" << j << ";\n";
        }
        o << "  {\n    const double opp = " << (pl == 1 ? "pcv * q2_" : "pcv * q1_") << r << ";\n"
          << "    const double own = " << (pl == 1 ? "q1_" : "q2_") << r << ";\n"
          << "    const double vp = " << (pl == 1 ? "x" + std::to_string(j) : (zs ? "-x" + std::to_string(j) : "y" + std::to_string(j))) << ";\n";
        for (int k = 0; k < nc; ++k) {
            std::string va = pl == 1 ? "x" + std::to_string(fc + k) : (zs ? "-x" + std::to_string(fc + k) : "y" + std::to_string(fc + k));
            o << "    contrib_r[" << k << " * ND + so" << r << "] = w * (" << va << " - vp) * opp;\n";
        }
        o << "    contrib_o[so" << r << "] = w * own;\n  }\n";
    }
    o << "  portal_v1[pi] = x0;\n";
    if (!zs) o << "  portal_v2[pi] = y0;\n";
    o << (tl ? "  tl_end(iter, tls);\n}\n" : "}\n");
    return o.str();
}

}  // namespace

namespace rlp {

// Compile `src` for sm_100a and load it; *lib_out stays null only under RLP_NO_JIT=1 (or in planning mode).
static int compile_and_load(const std::string &src, cudaLibrary_t *lib_out, std::string *log_out,
                            const char *program_name = "micro_k.cu") {
    *lib_out = nullptr;
    // RLP_NO_JIT=1 is the ONLY way onto the interpreter kernels for a tree whose shapes qualify for straight-line
    // code: a missing libnvrtc or a failed compile is an error, not a silent 3x slowdown.
    const char *no_jit = getenv("RLP_NO_JIT");
    if (no_jit && no_jit[0] == '1') return RLP_OK;
    if (!g_nvrtc.load())
        return fail(RLP_ERR_UNSUPPORTED, "libnvrtc.so.12 could not be loaded (set RLP_NO_JIT=1 to run the interpreter kernels)");
    nvrtcProgram prog;
    if (g_nvrtc.create(&prog, src.c_str(), program_name, 0, nullptr, nullptr) != NVRTC_SUCCESS)
        return fail(RLP_ERR_CUDA, "nvrtcCreateProgram failed");
    const char *opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17"};
    nvrtcResult res = g_nvrtc.compile(prog, 4, opts);
    if (res != NVRTC_SUCCESS) {
        size_t n = 0;
        g_nvrtc.log_size(prog, &n);
        std::string log(n, '\0');
        if (n) g_nvrtc.log(prog, &log[0]);
        if (log_out) *log_out = log;
        rlp::g_jit_failed++;
        g_nvrtc.destroy(&prog);
        return fail(RLP_ERR_CUDA, "NVRTC compile failed: %.400s", log.c_str());
    }
    rlp::g_jit_compiled++;
    if (rlp::g_dry_run) {                        // planning mode: the source compiled; nothing to load without a GPU
        g_nvrtc.destroy(&prog);
        return RLP_OK;
    }
    size_t sz = 0;
    g_nvrtc.cubin_size(prog, &sz);
    std::vector<char> cubin(sz);
    g_nvrtc.cubin(prog, cubin.data());
    g_nvrtc.destroy(&prog);
    cudaLibrary_t lib;
    cudaError_t e = cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RLP_ERR_CUDA, "cudaLibraryLoadData of a generated kernel: %s", cudaGetErrorString(e));
    }
    *lib_out = lib;
    return RLP_OK;
}

// Compile the straight-line kernel of `tm`; returns RLP_OK with *kernel == nullptr when JIT is unavailable.
// RLP_FUSED_DUMP=<path>: the source is also written to <path> and compiled under that name, so that a profiler's
// source view (ncu --import-source on, -lineinfo) can find it.
int compile_and_load_source(const std::string &src, cudaLibrary_t *lib_out, std::string *log_out) {
    const char *dump = getenv("RLP_FUSED_DUMP");
    if (dump && dump[0]) {
        if (FILE *f = fopen(dump, "w")) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
        return compile_and_load(src, lib_out, log_out, dump);
    }
    return compile_and_load(src, lib_out, log_out, "cfr_fused_k.cu");
}

int jit_micro_kernel(const MicroTmpl &tm, bool zs, cudaLibrary_t *lib_out, cudaKernel_t *kernel, std::string *log_out) {
    *kernel = nullptr;
    int rc = compile_and_load(generate(tm, zs), lib_out, log_out);
    if (rc || !*lib_out) return rc;
    if (cudaLibraryGetKernel(kernel, *lib_out, "micro_k") != cudaSuccess) {
        cudaGetLastError();
        cudaLibraryUnload(*lib_out);
        *lib_out = nullptr;
        *kernel = nullptr;
        return fail(RLP_ERR_CUDA, "generated library has no micro_k");
    }
    return RLP_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Upper tiles of one shape (round-1 betting nodes + board-deal chance nodes of a deal, micro roots as portal
// leaves): one thread per tile, straight-line.  Phase 0 pushes strategy reach down to the portals; phase 1 redoes
// that (a handful of multiplies), pulls the portal values up through the chance and decision nodes and scatters the
// contributions of the tile's decision nodes.
static std::string generate_upper(const UpperShape &sh, bool zs) {
    std::ostringstream o;
    const bool tl = want_timeline();
    const int n = (int)sh.kind.size();
    std::vector<int> dec_idx(n, -1), term_idx(n, -1), portal_idx(n, -1), edge_idx(n, -1), parent(n, -1), slot(n, 0);
    int nd = 0, nt = 0, np = 0, ne = 0;
    for (int j = 0; j < n; ++j) {
        if (sh.kind[j] == 1 || sh.kind[j] == 2) dec_idx[j] = nd++;
        else if (sh.kind[j] == -1) term_idx[j] = nt++;
        else if (sh.kind[j] == 3) { portal_idx[j] = sh.portal_rank[j]; ++np; }
        for (int k = 0; k < sh.nchild[j]; ++k) {
            int c = sh.first_child[j] + k;
            parent[c] = j; slot[c] = k;
            if (sh.kind[j] == 0) edge_idx[c] = ne++;
        }
    }
    auto S = [](const char *p, int j) { return std::string(p) + std::to_string(j); };
    o << "struct IterState { long long t; int linear; int pad; unsigned long long touched; unsigned long long* tl; unsigned int tl_count, tl_cap; };\n"
         "__device__ __forceinline__ unsigned long long rlp_now() { unsigned long long t; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(t)); return t; }\n"
         "__device__ __forceinline__ int tl_begin(const IterState* it, int kind) {\n"
         "  if (!it->tl || blockIdx.x != 0 || threadIdx.x != 0) return -1;\n"
         "  IterState* m = const_cast<IterState*>(it); unsigned int slot = atomicAdd(&m->tl_count, 1u);\n"
         "  if (slot >= m->tl_cap) return -1; m->tl[3 * slot] = (unsigned long long)kind; m->tl[3 * slot + 1] = rlp_now(); return (int)slot; }\n"
         "__device__ __forceinline__ void tl_end(const IterState* it, int slot) { if (slot >= 0) const_cast<IterState*>(it)->tl[3 * slot + 2] = rlp_now(); }\n";
    for (int phase = 0; phase < 2; ++phase) {
        o << "extern \"C\" __global__ void __launch_bounds__(64) upper_k" << phase << "(\n"
             "    const int* __restrict__ colbase, const int* __restrict__ slot_r, const int* __restrict__ slot_o,\n"
             "    const double* __restrict__ pcs, const double* __restrict__ cprob, const double* __restrict__ pay1,\n"
             "    const double* __restrict__ pay2, const long long* __restrict__ poff, const double* __restrict__ sigma,\n"
             "    double* __restrict__ portal_p1, double* __restrict__ portal_p2, const double* __restrict__ portal_v1,\n"
             "    const double* __restrict__ portal_v2, double* __restrict__ contrib_r, double* __restrict__ contrib_o,\n"
             "    double* __restrict__ root_val, const IterState* __restrict__ iter, long long T,\n"
             "    const double* __restrict__ fan_v1, const double* __restrict__ fan_v2, long long ND) {\n"
             "  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;\n"
             "  asm volatile(\"griddepcontrol.launch_dependents;\" ::: \"memory\");\n"
             "  if (i >= T) { asm volatile(\"griddepcontrol.wait;\" ::: \"memory\"); return; }\n"
             "  (void)poff;\n";
        for (int j = 0; j < n; ++j)
            if (dec_idx[j] >= 0) {
                int d = dec_idx[j];
                o << "  const int cb" << j << " = colbase[" << d << " * T + i];\n";
                if (phase == 1)
                    o << "  const int so" << j << " = slot_o[" << d << " * T + i];\n"
                      << "  const double pc" << j << " = pcs[" << d << " * T + i];\n";
            }
        if (phase == 1) {
            for (int j = 0; j < n; ++j) {
                if (edge_idx[j] >= 0) o << "  const double e" << j << " = cprob[" << edge_idx[j] << " * T + i];\n";
                if (term_idx[j] >= 0) {
                    o << "  const double x" << j << " = pay1[" << term_idx[j] << " * T + i];\n";
                    if (!zs) o << "  const double y" << j << " = pay2[" << term_idx[j] << " * T + i];\n";
                }
            }
        }
        o << "  asm volatile(\"griddepcontrol.wait;\" ::: \"memory\");\n"
          << (tl ? (phase == 0 ? "  const int tls = tl_begin(iter, 0);\n" : "  const int tls = tl_begin(iter, 2);\n") : "");
        if (phase == 1) o << "  const double w = iter->linear ? (double)iter->t : 1.0;\n";
        for (int j = 1; j < n; ++j)
            if (parent[j] >= 0 && dec_idx[parent[j]] >= 0)
                o << "  const double s" << j << " = sigma[cb" << parent[j] << " + " << slot[j] << "];\n";
        if (phase == 1)
            for (int j = 0; j < n; ++j)
                if (sh.kind[j] == 4) {
                    o << "  const double x" << j << " = fan_v1[" << sh.fan_index[j] << " * T + i];\n";
                    if (!zs) o << "  const double y" << j << " = fan_v2[" << sh.fan_index[j] << " * T + i];\n";
                }
        if (phase == 1)
            for (int j = 0; j < n; ++j)
                if (portal_idx[j] >= 0) {
                    o << "  const double x" << j << " = portal_v1[" << portal_idx[j] << " * T + i];\n";
                    if (!zs) o << "  const double y" << j << " = portal_v2[" << portal_idx[j] << " * T + i];\n";
                }
        // forward
        o << "  const double q1_0 = 1.0, q2_0 = 1.0;\n";
        for (int j = 0; j < n; ++j) {
            int kd = sh.kind[j];
            if (kd == -1 || kd == 3 || kd == 4) continue;
            for (int k = 0; k < sh.nchild[j]; ++k) {
                int c = sh.first_child[j] + k;
                if (sh.kind[c] == -1) continue;
                std::string a = (kd == 1 ? S("s", c) + " * " : std::string()) + S("q1_", j);
                std::string b = (kd == 2 ? S("s", c) + " * " : std::string()) + S("q2_", j);
                if (sh.kind[c] == 4) {          // chance node whose children are all portals: they inherit its reach
                    if (phase == 0)
                        o << "  portal_p1[" << sh.fan_k0[c] << " * T + i] = " << a << ";\n  portal_p2[" << sh.fan_k0[c]
                          << " * T + i] = " << b << ";\n";
                } else if (sh.kind[c] == 3) {
                    if (phase == 0)
                        o << "  portal_p1[" << portal_idx[c] << " * T + i] = " << a << ";\n  portal_p2[" << portal_idx[c]
                          << " * T + i] = " << b << ";\n";
                } else {
                    o << "  const double q1_" << c << " = " << a << ";\n  const double q2_" << c << " = " << b << ";\n";
                }
            }
        }
        if (phase == 0) { o << (tl ? "  tl_end(iter, tls);\n}\n" : "}\n"); continue; }
        // backward
        for (int j = n - 1; j >= 0; --j) {
            int kd = sh.kind[j];
            if (kd == -1 || kd == 3 || kd == 4) continue;
            int fc = sh.first_child[j], nc = sh.nchild[j];
            o << "  double a" << j << " = 0.0;\n";
            for (int k = 0; k < nc; ++k)
                o << "  a" << j << " = a" << j << " + " << (kd == 0 ? S("e", fc + k) : S("s", fc + k)) << " * x" << (fc + k) << ";\n";
            o << "  const double x" << j << " = a" << j << ";\n";
            if (!zs) {
                o << "  double b" << j << " = 0.0;\n";
                for (int k = 0; k < nc; ++k)
                    o << "  b" << j << " = b" << j << " + " << (kd == 0 ? S("e", fc + k) : S("s", fc + k)) << " * y" << (fc + k) << ";\n";
                o << "  const double y" << j << " = b" << j << ";\n";
            }
            if (kd == 0) continue;
            o << "  {\n    const double opp = pc" << j << (kd == 1 ? " * q2_" : " * q1_") << j << ";\n"
              << "    const double own = " << (kd == 1 ? "q1_" : "q2_") << j << ";\n"
              << "    const double vp = " << (kd == 1 ? S("x", j) : (zs ? "-" + S("x", j) : S("y", j))) << ";\n";
            for (int k = 0; k < nc; ++k) {
                std::string va = kd == 1 ? S("x", fc + k) : (zs ? "-" + S("x", fc + k) : S("y", fc + k));
                o << "    contrib_r[" << k << " * ND + so" << j << "] = w * (" << va << " - vp) * opp;\n";
            }
            o << "    contrib_o[so" << j << "] = w * own;\n  }\n";
        }
        o << "  root_val[i] = x0;\n" << (tl ? "  tl_end(iter, tls);\n}\n" : "}\n");
    }
    return o.str();
}

int jit_upper_kernels(const UpperShape &sh, bool zs, cudaLibrary_t *lib_out, cudaKernel_t *k0, cudaKernel_t *k1, std::string *log_out) {
    *k0 = *k1 = nullptr;
    int rc = compile_and_load(generate_upper(sh, zs), lib_out, log_out);
    if (rc || !*lib_out) return rc;
    if (cudaLibraryGetKernel(k0, *lib_out, "upper_k0") != cudaSuccess || cudaLibraryGetKernel(k1, *lib_out, "upper_k1") != cudaSuccess) {
        cudaGetLastError();
        cudaLibraryUnload(*lib_out);
        *lib_out = nullptr;
        *k0 = *k1 = nullptr;
        return fail(RLP_ERR_CUDA, "generated library has no upper_k0 / upper_k1");
    }
    return RLP_OK;
}

std::string jit_micro_source(const MicroTmpl &tm, bool zs) { return generate(tm, zs); }

}  // namespace rlp
