// b200ens.cpp -- host runtime of libb200ens.so (C ABI in include/b200ens.h).
//
//  * b200ens_compile : (model CUDA-C source x stepper kernel x dtype) -> NVRTC -> sm_100a cubin
//  * b200ens_solve   : shard trajectory ranges over the GPUs of the box, one host thread and
//                      two streams per device, chunked H2D / kernel / D2H pipeline, host gather
//  * b200ens_solve_device : one launch on caller-owned device buffers and stream
//
// There is no CPU fallback anywhere in this file: without a CUDA device the solve entry
// points return B200ENS_E_NODEVICE.
#include "../../include/b200ens.h"

#include <cuda_runtime.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct EmbeddedHeader {
    const char* name;
    const char* text;
};
const EmbeddedHeader kHeaders[] = {
#include "kernels_embed.inc"
};
constexpr int kNumHeaders = sizeof(kHeaders) / sizeof(kHeaders[0]);

thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                    \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(e_ == cudaErrorMemoryAllocation ? B200ENS_E_NOMEM : B200ENS_E_CUDA,         \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Mirror of struct B2Args in kernels/b2_common.cuh (same member order and types).
struct B2Args {
    const void* u0;
    const void* p;
    const void* saveat;
    const void* dW;
    void* out_u;
    int* retcode;
    b200ens_stats* stats;
    unsigned long long* work_counter;
    long long N;
    long long maxiters;
    long long nsteps_noise;
    unsigned long long seed;
    unsigned long long traj_offset;
    double t0, t1, dt, abstol, reltol, dtmin, dtmax, qmin, qmax, gamma, beta1, beta2, qoldinit;
    int n_save, adaptive, refill_threshold, stage_stride;
    int noise_injected, event_terminate, interp_points, save_tstops;
    float f_t0, f_t1, f_dt, f_abstol, f_reltol, f_dtmin, f_dtmax, f_qmin, f_qmax, f_gamma, f_beta1, f_beta2, f_qoldinit;
    int pad0_;
    const unsigned* perm;
    double tol_a[32], tol_r[32];
    float f_tol_a[32], f_tol_r[32];
    void* every_t;
    int save_every, pad1_;
    double* mom_sum;
    double* mom_sq;
    unsigned long long* mom_fail;
    const void* tstops;
    int n_tstops, pad2_;
};

bool is_sde(int alg) { return alg == B200ENS_EM || alg == B200ENS_SOSRA || alg == B200ENS_SRIW1; }
bool is_rosenbrock(int alg) {
    return alg == B200ENS_ROSENBROCK23 || alg == B200ENS_RODAS4 || alg == B200ENS_RODAS5 || alg == B200ENS_RODAS5P;
}
bool needs_jac(int alg) { return is_rosenbrock(alg) || alg == B200ENS_FBDF; }   // W = I/(gamma dt) - J or I - beta dt J
int alg_order(int alg) {
    switch (alg) {
    case B200ENS_TSIT5: return 5;
    case B200ENS_VERN7: return 7;
    case B200ENS_ROSENBROCK23: return 2;
    case B200ENS_RODAS4: return 4;
    case B200ENS_RODAS5:
    case B200ENS_RODAS5P: return 5;
    default: return 1;
    }
}
double dflt(double v, double d) { return (v != v || v < 0) ? d : v; }

constexpr int kBlock = 128;

}  // namespace

struct b200ens_model {
    int n_state = 0, n_param = 0, dtype = 0, alg = 0;
    int n_out = 0;          // entries of an output row: n_state, or the number of save_idxs
    unsigned flags = 0;
    bool has_event = false, has_noise = false;
    std::string name, source, log;
    std::vector<char> cubin;
    int regs = -1, smem = -1, lmem = -1, spill = 0, min_blocks = 1;
    int block = 128;        // threads per CTA the kernel was compiled for (B2_BLOCK)
    int ksmem = 0;          // 1: ERK stage vectors in shared memory
    int split = 0;          // 1: one trajectory per lane of a 4-warp CTA, components split over the warps (b2_split.cuh)
    int kvec_bytes = 0;     // shared-memory bytes per thread for them
    int has_mass = 0;       // the RHS source carries a constant mass matrix (B2_HAS_MASS)
    std::mutex mu;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
    cudaKernel_t k_work_keys = nullptr, k_work_scatter = nullptr;   // expected-work ordering (kernels/b2_work.cuh)
    cudaKernel_t kernel_adaptive = nullptr;   // specialised entry (adaptive=1, save_tstops=0) when the module has one
    size_t elem() const { return dtype == B200ENS_F64 ? 8 : 4; }
};

namespace {

// ---------------------------------------------------------------- NVRTC, loaded explicitly
// The JIT compiler is dlopen'ed by PATH instead of being bound at link time: a process that imported PyTorch first
// already holds torch's bundled libnvrtc.so.12 (12.8 in this image), and link-time binding would silently pick that
// one up instead of the CUDA toolkit's (12.9) -- different ptxas, different register allocation.  Search order: $B200ENS_NVRTC, the toolkit the library was built against, then the default loader path.
struct NvrtcApi {
    void* handle = nullptr;
    int major = 0, minor = 0;
    std::string path;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    const char* (*GetErrorString)(nvrtcResult) = nullptr;
    nvrtcResult (*Version)(int*, int*) = nullptr;
};
#ifndef B200ENS_NVRTC_DEFAULT
#define B200ENS_NVRTC_DEFAULT "/usr/local/cuda/lib64/libnvrtc.so.12"
#endif
NvrtcApi* nvrtc_api() {
    static NvrtcApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        std::vector<std::string> cand;
        if (const char* e = getenv("B200ENS_NVRTC")) cand.push_back(e);
        cand.push_back(B200ENS_NVRTC_DEFAULT);
        cand.push_back("libnvrtc.so.12");
        cand.push_back("libnvrtc.so");
        for (const auto& c : cand) {
            // DEEPBIND: the library must resolve its own internal symbols, not those of another NVRTC already in the
            // global scope (torch preloads its bundled 12.8 with RTLD_GLOBAL)
            void* h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
            if (!h) continue;
            NvrtcApi a;
            a.handle = h;
            a.path = c;
#define B2_SYM(field, name) *(void**)(&a.field) = dlsym(h, name)
            B2_SYM(CreateProgram, "nvrtcCreateProgram");
            B2_SYM(CompileProgram, "nvrtcCompileProgram");
            B2_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
            B2_SYM(GetProgramLog, "nvrtcGetProgramLog");
            B2_SYM(GetCUBINSize, "nvrtcGetCUBINSize");
            B2_SYM(GetCUBIN, "nvrtcGetCUBIN");
            B2_SYM(DestroyProgram, "nvrtcDestroyProgram");
            B2_SYM(GetErrorString, "nvrtcGetErrorString");
            B2_SYM(Version, "nvrtcVersion");
#undef B2_SYM
            if (!a.CreateProgram || !a.CompileProgram || !a.GetProgramLogSize || !a.GetProgramLog || !a.GetCUBINSize ||
                !a.GetCUBIN || !a.DestroyProgram || !a.GetErrorString || !a.Version) {
                dlclose(h);
                continue;
            }
            a.Version(&a.major, &a.minor);
            api = a;
            return;
        }
    });
    return api.handle ? &api : nullptr;
}
#define B2_NVRTC_OR_FAIL(var)                                                                                   \
    NvrtcApi* var = nvrtc_api();                                                                                \
    if (!var) return fail(B200ENS_E_COMPILE, "NVRTC not found (set B200ENS_NVRTC to the path of libnvrtc.so.12)")

// ---------------------------------------------------------------- compile cache (in-process; the on-disk level follows)
std::mutex g_cache_mu;
std::map<std::string, std::shared_ptr<std::pair<std::vector<char>, std::string>>> g_cache;

// Registers / stack frame / spills of the ensemble entry points (max over the generic and the specialised entry; the
// work-ordering helper kernels of the same module are ignored).  Authoritative source: the cubin itself (.nv.info
// EIATTR_REGCOUNT / EIATTR_FRAME_SIZE).  The ptxas -v log only refines the spill figure: NVRTC 12.9 answers repeated
// compilations of the same source from an on-disk cache with an EMPTY log (measured: every process after the first),
// so nothing here may depend on the log being present.
bool cubin_resources(const std::vector<char>& cubin, int* regs, int* frame) {
    const unsigned char* d = (const unsigned char*)cubin.data();
    const size_t n = cubin.size();
    if (n < 0x40 || memcmp(d, "\177ELF", 4) != 0 || d[4] != 2) return false;
    auto rd = [&](size_t off, int bytes) -> unsigned long long {
        unsigned long long v = 0;
        if (off + bytes > n) return 0;
        memcpy(&v, d + off, bytes);
        return v;
    };
    const size_t shoff = rd(0x28, 8);
    const int shentsize = (int)rd(0x3A, 2), shnum = (int)rd(0x3C, 2), shstrndx = (int)rd(0x3E, 2);
    if (!shoff || shentsize < 64 || shstrndx >= shnum) return false;
    auto sh = [&](int i, int field_off, int bytes) { return rd(shoff + (size_t)i * shentsize + field_off, bytes); };
    const size_t shstr = sh(shstrndx, 0x18, 8);
    int i_info = -1, i_sym = -1;
    for (int i = 0; i < shnum; i++) {
        const char* nm = (const char*)d + shstr + sh(i, 0, 4);
        if ((size_t)(nm - (const char*)d) >= n) continue;
        if (!strcmp(nm, ".nv.info")) i_info = i;
        if (sh(i, 4, 4) == 2) i_sym = i;   // SHT_SYMTAB
    }
    if (i_info < 0 || i_sym < 0) return false;
    const size_t symoff = sh(i_sym, 0x18, 8), stroff = sh((int)sh(i_sym, 0x28, 4), 0x18, 8);
    auto is_ens = [&](unsigned symidx) {
        const size_t name_off = stroff + rd(symoff + (size_t)symidx * 24, 4);
        return name_off < n && !strncmp((const char*)d + name_off, "b2_ensemble_kernel", strlen("b2_ensemble_kernel"));
    };
    size_t p = sh(i_info, 0x18, 8);
    const size_t end = p + sh(i_info, 0x20, 8);
    int r = -1, f = 0;
    while (p + 4 <= end && end <= n) {
        const int fmt = d[p], attr = d[p + 1], size = (int)rd(p + 2, 2);
        p += 4;
        if (fmt != 4) continue;   // NVAL/BVAL/HVAL carry their value in the size field
        if (size >= 8 && (attr == 0x2f || attr == 0x11) && is_ens((unsigned)rd(p, 4))) {
            const int v = (int)rd(p + 4, 4);
            if (attr == 0x2f) r = std::max(r, v);   // EIATTR_REGCOUNT
            else f = std::max(f, v);                // EIATTR_FRAME_SIZE
        }
        p += size;
    }
    *regs = r;
    *frame = f;
    return r >= 0;
}

void parse_ptxas_log(b200ens_model* m) {
    m->regs = -1;
    m->smem = m->lmem = 0;
    m->spill = -1;
    if (getenv("B200ENS_IGNORE_PTXAS_LOG")) m->log.clear();   // tests: behave as if NVRTC served the compile from its cache
    const char* s = m->log.c_str();
    const char* pos = s;
    auto num_before = [&](const char* blk, const char* end, const char* tag) {
        const char* q = strstr(blk, tag);
        if (!q || q >= end) return -1;
        const char* b = q;
        while (b > blk && isdigit((unsigned char)b[-1])) b--;
        return atoi(b);
    };
    while ((pos = strstr(pos, "Compiling entry function '")) != nullptr) {
        const char* name = pos + strlen("Compiling entry function '");
        const char* next = strstr(name, "Compiling entry function '");
        const char* end = next ? next : s + m->log.size();
        if (!strncmp(name, "b2_ensemble_kernel", strlen("b2_ensemble_kernel"))) {
            const char* q = strstr(name, "Used ");
            if (q && q < end) m->regs = std::max(m->regs, atoi(q + 5));
            m->smem = std::max(m->smem, num_before(name, end, " bytes smem"));
            m->lmem = std::max(m->lmem, num_before(name, end, " bytes stack frame"));
            m->spill = std::max(m->spill, num_before(name, end, " bytes spill stores"));
        }
        pos = name;
    }
    int regs = -1, frame = 0;
    if (cubin_resources(m->cubin, &regs, &frame)) {
        m->regs = regs;
        m->lmem = frame;
    }
    // no log (compile served from NVRTC's cache): a stack frame of F bytes corresponds to ~2F bytes of spill stores
    if (m->spill < 0) m->spill = 2 * m->lmem;
}

std::string build_source(const b200ens_model_desc* d, int min_blocks, int block, int ksmem, int split = 0) {
    char head[1024];
    snprintf(head, sizeof head,
             "// generated by libb200ens (model '%s')\n"
             "#define B2_F64 %d\n#define B2_NSTATE %d\n#define B2_NPARAM %d\n#define B2_ALG %d\n"
             "#define B2_HAS_JAC %d\n#define B2_HAS_TGRAD %d\n#define B2_HAS_NOISE %d\n#define B2_HAS_EVENT %d\n#define B2_HAS_DEVENT %d\n"
             "#define B2_BLOCK %d\n#define B2_MINBLOCKS %d\n#define B2_KSMEM %d\n#define B2_SPLIT %d\n#include \"b2_common.cuh\"\n",
             d->name ? d->name : "", d->dtype == B200ENS_F64 ? 1 : 0, d->n_state, d->n_param, d->alg,
             d->jac_src ? 1 : 0, d->tgrad_src ? 1 : 0, d->noise_src ? 1 : 0,
             (d->condition_src && d->affect_src) ? 1 : 0, (d->dcondition_src && d->daffect_src) ? 1 : 0, block, min_blocks,
             ksmem, split);
    std::string s;
    // B200ENS_DEFINES="NAME=VALUE,NAME2=VALUE2": extra #defines in front of everything (experiments and A/B tests of
    // kernel variants, e.g. B2_PACK2=0); part of the source, hence of every cache key
    if (const char* e = getenv("B200ENS_DEFINES")) {
        std::string defs = e;
        size_t pos = 0;
        while (pos < defs.size()) {
            size_t end = defs.find(',', pos);
            if (end == std::string::npos) end = defs.size();
            std::string item = defs.substr(pos, end - pos);
            const size_t eq = item.find('=');
            if (!item.empty()) s += "#define " + (eq == std::string::npos ? item : item.substr(0, eq) + " " + item.substr(eq + 1)) + "\n";
            pos = end + 1;
        }
    }
    // (before the prelude: b2_common.cuh reads B2_NOUT)
    if (d->n_save_idxs > 0) {   // solve(...; save_idxs): B2_NOUT + the component of every output column (b2_common.cuh)
        s += "#define B2_NOUT " + std::to_string(d->n_save_idxs) + "\nstatic constexpr int B2_SAVE_IDXS_[B2_NOUT] = {";
        for (int i = 0; i < d->n_save_idxs; i++) s += (i ? ", " : "") + std::to_string(d->save_idxs[i]);
        s += "};\n";
    }
    s += head;
    for (const char* part : {d->rhs_src, d->jac_src, d->tgrad_src, d->noise_src, d->condition_src, d->affect_src,
                             d->dcondition_src, d->daffect_src})
        if (part) {
            s += part;
            s += "\n";
        }
    if (d->flags & B200ENS_MODEL_SDE_ADAPTIVE) s += "#define B2_SDE_ADAPT 1\n";
    s += "#include \"b2_entry.cuh\"\n";
    return s;
}

// ---------------------------------------------------------------- compile cache (on disk)
// A model is compiled once per (generated source, embedded kernel headers, flags, NVRTC version): the cubin and the
// ptxas log are kept under $B200ENS_CACHE_DIR, else <directory of libb200ens.so>/cubin_cache (in-tree: it travels with
// the library), else ~/.cache/b200ens.  Config 5's split Vern7 kernel takes 14.5 s to JIT (profiles/README.md); every
// later process loads it in milliseconds.  B200ENS_CACHE=0 switches the disk cache off.  Entries are written to a
// temporary file and renamed, so concurrent processes (one per GPU under torchrun) never see a partial cubin.
uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; i++) {
        h ^= p[i];
        h *= 0x100000001b3ull;
    }
    return h;
}
const std::string& cache_dir() {
    static std::string dir;
    static std::once_flag once;
    std::call_once(once, [] {
        if (const char* e = getenv("B200ENS_CACHE")) if (atoi(e) == 0) return;
        std::vector<std::string> cand;
        if (const char* e = getenv("B200ENS_CACHE_DIR")) cand.push_back(e);
        Dl_info info;
        if (dladdr((void*)&fnv1a, &info) && info.dli_fname) {
            std::string so = info.dli_fname;
            const size_t sl = so.rfind('/');
            cand.push_back((sl == std::string::npos ? std::string(".") : so.substr(0, sl)) + "/cubin_cache");
        }
        if (const char* h = getenv("HOME")) cand.push_back(std::string(h) + "/.cache/b200ens");
        for (const auto& c : cand) {
            std::string cmd;
            // mkdir -p by hand (two levels are enough for the candidates above)
            for (size_t i = 1; i <= c.size(); i++)
                if (i == c.size() || c[i] == '/') (void)::mkdir(c.substr(0, i).c_str(), 0755);
            const std::string probe = c + "/.probe" + std::to_string((long long)getpid());
            FILE* f = fopen(probe.c_str(), "wb");
            if (!f) continue;
            fclose(f);
            remove(probe.c_str());
            dir = c;
            return;
        }
    });
    return dir;
}
std::string cache_key(const b200ens_model* m, bool fast, const NvrtcApi* nv) {
    static uint64_t hdr_a = 0, hdr_b = 0;
    static std::once_flag once;
    std::call_once(once, [] {
        hdr_a = 0xcbf29ce484222325ull;
        hdr_b = 0x84222325cbf29ce4ull;
        for (int i = 0; i < kNumHeaders; i++) {
            hdr_a = fnv1a(kHeaders[i].text, strlen(kHeaders[i].text), fnv1a(kHeaders[i].name, strlen(kHeaders[i].name), hdr_a));
            hdr_b = fnv1a(kHeaders[i].text, strlen(kHeaders[i].text), hdr_b * 31 + 7);
        }
    });
    char tail[96];
    snprintf(tail, sizeof tail, "|%s|nvrtc%d.%d|sm_100a|v1", fast ? "fmad" : "nofmad", nv ? nv->major : 0, nv ? nv->minor : 0);
    uint64_t a = fnv1a(m->source.data(), m->source.size(), hdr_a), b = fnv1a(m->source.data(), m->source.size(), hdr_b);
    a = fnv1a(tail, strlen(tail), a);
    b = fnv1a(tail, strlen(tail), b);
    char name[64];
    snprintf(name, sizeof name, "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
    return name;
}
bool read_file(const std::string& path, std::vector<char>* out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out->resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && fread(out->data(), 1, out->size(), f) == out->size();
    fclose(f);
    return ok;
}
void write_file_atomic(const std::string& path, const void* data, size_t n) {
    const std::string tmp = path + ".tmp" + std::to_string((long long)getpid());
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data, 1, n, f) == n;
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
}

int nvrtc_compile(b200ens_model* m) {
    const bool fast = (m->flags & B200ENS_MODEL_FAST_MATH) != 0;
    const std::string key = std::string(fast ? "F" : "S") + m->source;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_cache.find(key);
        if (it != g_cache.end()) {
            m->cubin = it->second->first;
            m->log = it->second->second;
            return 0;
        }
    }
    B2_NVRTC_OR_FAIL(nv);
    if (const char* dd = getenv("B200ENS_DUMP_SRC")) {   // tools: write the generated translation unit for offline nvcc / cuobjdump
        static std::atomic<int> seq{0};
        const std::string path = std::string(dd) + "/b200ens_model_" + std::to_string(seq.fetch_add(1)) + ".cu";
        write_file_atomic(path, m->source.data(), m->source.size());
    }
    const std::string& cdir = cache_dir();
    const std::string cfile = cdir.empty() ? std::string() : cdir + "/" + cache_key(m, fast, nv);
    if (!cfile.empty()) {
        std::vector<char> cub, lg;
        if (read_file(cfile + ".cubin", &cub) && cub.size() > 64 && memcmp(cub.data(), "\177ELF", 4) == 0) {
            read_file(cfile + ".log", &lg);
            m->cubin = cub;
            m->log.assign(lg.begin(), lg.end());
            std::lock_guard<std::mutex> lk(g_cache_mu);
            g_cache[key] = std::make_shared<std::pair<std::vector<char>, std::string>>(m->cubin, m->log);
            return 0;
        }
    }
    nvrtcProgram prog;
    const char* hdr_names[kNumHeaders];
    const char* hdr_text[kNumHeaders];
    for (int i = 0; i < kNumHeaders; i++) {
        hdr_names[i] = kHeaders[i].name;
        hdr_text[i] = kHeaders[i].text;
    }
    nvrtcResult r = nv->CreateProgram(&prog, m->source.c_str(), "b200ens_model.cu", kNumHeaders, hdr_text, hdr_names);
    if (r != NVRTC_SUCCESS) return fail(B200ENS_E_COMPILE, "nvrtcCreateProgram: %s", nv->GetErrorString(r));
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "--ptxas-options=-v",
                                     fast ? "--fmad=true" : "--fmad=false", "--prec-div=true", "--prec-sqrt=true",
                                     "--ftz=false", "-default-device"};   // -default-device: generic lambdas inside device code (b2_sde.cuh)
    r = nv->CompileProgram(prog, (int)opts.size(), opts.data());
    size_t logn = 0;
    nv->GetProgramLogSize(prog, &logn);
    m->log.assign(logn, '\0');
    if (logn) nv->GetProgramLog(prog, &m->log[0]);
    if (r != NVRTC_SUCCESS) {
        nv->DestroyProgram(&prog);
        return fail(B200ENS_E_COMPILE, "NVRTC compile failed (%s):\n%.1500s", nv->GetErrorString(r), m->log.c_str());
    }
    size_t n = 0;
    if (nv->GetCUBINSize(prog, &n) != NVRTC_SUCCESS || n == 0) {
        nv->DestroyProgram(&prog);
        return fail(B200ENS_E_COMPILE, "NVRTC produced no cubin");
    }
    m->cubin.resize(n);
    nv->GetCUBIN(prog, m->cubin.data());
    nv->DestroyProgram(&prog);
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_cache[key] = std::make_shared<std::pair<std::vector<char>, std::string>>(m->cubin, m->log);
    }
    if (!cfile.empty()) {
        write_file_atomic(cfile + ".log", m->log.data(), m->log.size());
        write_file_atomic(cfile + ".cubin", m->cubin.data(), m->cubin.size());
    }
    return 0;
}

int ensure_loaded(b200ens_model* m) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->kernel) return 0;
    CU(cudaLibraryLoadData(&m->lib, m->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    CU(cudaLibraryGetKernel(&m->kernel, m->lib, "b2_ensemble_kernel"));
    // optional entry points are looked up only in modules that define them (kernels/b2_entry.cuh): no failing API
    // calls on the normal path (they show up as errors under compute-sanitizer)
    const bool erk = m->alg == B200ENS_TSIT5 || m->alg == B200ENS_VERN7;
    if (erk || needs_jac(m->alg))
        CU(cudaLibraryGetKernel(&m->kernel_adaptive, m->lib, "b2_ensemble_kernel_adaptive"));
    if (!is_sde(m->alg)) {
        CU(cudaLibraryGetKernel(&m->k_work_keys, m->lib, "b2_work_keys"));
        CU(cudaLibraryGetKernel(&m->k_work_scatter, m->lib, "b2_work_scatter"));
    }
    return 0;
}

// ---------------------------------------------------------------- per-device state
struct Slot {  // one pipeline slot = one stream + device buffers for one chunk
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // h2d start, kernel start, kernel end, d2h end
    void *u0 = nullptr, *p = nullptr, *out = nullptr, *dW = nullptr;
    int* rc = nullptr;
    b200ens_stats* stats = nullptr;
    unsigned long long* counter = nullptr;
    void* work = nullptr;   // expected-work ordering scratch (counter | histogram | cursors | perm | keys)
    size_t cap_work = 0;
    void* every_t = nullptr;   // save_everystep: step times [chunk][capacity]
    size_t cap_every = 0;
    size_t cap_u0 = 0, cap_p = 0, cap_out = 0, cap_dW = 0, cap_n = 0;
    // pinned bounce buffers for callers whose arrays are pageable (e.g. plain Julia Arrays)
    char *h_in = nullptr, *h_out = nullptr;
    size_t cap_hin = 0, cap_hout = 0;
    // fused ensemble moments: this chunk's [sum | sumsq | failures] on the device and its pinned host copy
    void* macc = nullptr;
    size_t cap_macc = 0;
    char* h_macc = nullptr;
    size_t cap_hmacc = 0;
};
constexpr int kMaxSlots = 4;
struct DeviceCtx {
    int dev = -1;
    int sms = 0;
    size_t mem_total = 0;
    std::mutex mu;  // one host-buffer solve at a time per device
    Slot slot[kMaxSlots];
    void* saveat = nullptr;
    size_t cap_save = 0;
    void* tstops = nullptr;   // solve(...; tstops) in the state type
    size_t cap_tstops = 0;
    void* acc = nullptr;      // ensemble-moments accumulators
    size_t cap_acc = 0;
    unsigned long long* counters = nullptr;  // ring for solve_device
    std::atomic<unsigned> ring{0};
    bool init = false;
};
constexpr int kMaxDev = 32;
constexpr int kRing = 256;
DeviceCtx g_dev[kMaxDev];
std::mutex g_dev_mu;

int device_ctx(int dev, DeviceCtx** out) {
    if (dev < 0 || dev >= kMaxDev) return fail(B200ENS_E_INVALID, "device %d out of range", dev);
    DeviceCtx& d = g_dev[dev];
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!d.init) {
        CU(cudaSetDevice(dev));
        d.dev = dev;
        CU(cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev));
        {
            size_t fr = 0;
            CU(cudaMemGetInfo(&fr, &d.mem_total));
        }
        {   // keep stream-ordered scratch (work ordering in b200ens_solve_device) cached in the default pool
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long thr = ~0ull;
                (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
            (void)cudaGetLastError();
        }
        for (auto& s : d.slot) {
            CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
            for (auto& e : s.ev) CU(cudaEventCreate(&e));
            CU(cudaMalloc(&s.counter, sizeof(unsigned long long)));
        }
        CU(cudaMalloc(&d.counters, kRing * sizeof(unsigned long long)));
        d.init = true;
    }
    *out = &d;
    return 0;
}

int grow(void** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*ptr) CU(cudaFree(*ptr));
    *ptr = nullptr;
    *cap = 0;
    const size_t want = need + need / 8 + 256;
    CU(cudaMalloc(ptr, want));
    *cap = want;
    return 0;
}

// ---------------------------------------------------------------- pageable host buffers
bool is_pinned(const void* ptr) {
    if (!ptr) return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}
void par_memcpy(void* dst, const void* src, size_t bytes) {
    const size_t kMin = 2u << 20;
    int nt = (int)std::min<size_t>(8, bytes / kMin);
    if (nt <= 1) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = (bytes / nt + 63) & ~(size_t)63;
    for (int i = 0; i < nt; i++) {
        const size_t off = (size_t)i * per;
        if (off >= bytes) break;
        const size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy((char*)dst + off, (const char*)src + off, len); });
    }
    for (auto& t : th) t.join();
}
int grow_host(char** ptr, size_t* cap, size_t need) {
    if (need <= *cap) return 0;
    if (*ptr) CU(cudaFreeHost(*ptr));
    *ptr = nullptr;
    *cap = 0;
    CU(cudaHostAlloc((void**)ptr, need + need / 8 + 256, cudaHostAllocDefault));
    *cap = need + need / 8 + 256;
    return 0;
}

struct LaunchPlan {
    int grid = 0, block = kBlock, smem = 0, stride = 0, refill = 0;
};

int plan_launch(b200ens_model* m, const b200ens_opts* o, DeviceCtx* d, long long N, int n_save, LaunchPlan* lp) {
    const int block = m->block;
    const size_t es = m->elem();
    const size_t ksm = m->ksmem ? (size_t)m->kvec_bytes * block : 0;
    int stride = n_save * m->n_out;
    if (stride % 2 == 0) stride += 1;  // odd row stride: conflict-free staging rows
    size_t smem = (size_t)block * stride * es;
    int nb_direct = 0, nb_staged = 0;
    if (ksm > 48 * 1024)
        CU(cudaFuncSetAttribute((const void*)m->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_direct, (const void*)m->kernel, block, ksm));
    bool staged = false;
    // Measured on B200 (profiles/): for the saveat shapes of configs 1-4 direct global stores are ~3%
    // faster than shared-memory staging (L2 merges the 12-byte rows), so auto means direct.
    if (o->stage_outputs > 0 && n_save > 0 && !m->split && smem + ksm <= 200 * 1024) {
        if (smem + ksm > 48 * 1024)
            CU(cudaFuncSetAttribute((const void*)m->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + ksm)));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb_staged, (const void*)m->kernel, block, smem + ksm));
        // stage when it costs at most a quarter of the occupancy (or when forced)
        staged = nb_staged >= 1 && (o->stage_outputs > 0 || 4 * nb_staged >= 3 * nb_direct);
    }
    if (o->stage_outputs > 0 && !staged && !m->split)
        return fail(B200ENS_E_UNSUPPORTED, "stage_outputs=1 but %zu bytes of shared memory per block do not fit", smem);
    int nb = staged ? nb_staged : nb_direct;
    // stiff steppers beyond 10 states run their LU with rolled loops on a local-memory matrix on a local-memory matrix (b2_rosenbrock.cuh,
    // B2_LU_ROLLED): ~3.5 KB of local memory per thread.  One CTA (of 64 threads, b200ens_compile) per SM keeps it inside the
    // L1; more resident threads spill it to the L2 and are slower (profiles/README.md)
    const int nb_occ = nb;
    // resident threads per SM so that their local memory (~224 KB) stays inside the L1: measured best at n = 12 (2 KB per
    // thread): 128 threads (38 ms per 100k trajectories; 64: 64 ms, 256: 51 ms), at n = 16 (3.4 KB): 64 threads
    if (needs_jac(m->alg) && m->n_state > 10 && m->lmem >= 1024)
        nb = std::min(nb, std::max(1, (int)(229376.0 / ((double)m->lmem * block) + 0.5)));
    if (const char* e = getenv("B200ENS_BLOCKS_PER_SM")) nb = std::max(1, std::min(nb_occ, atoi(e)));  // experiments
    if (nb < 1) return fail(B200ENS_E_CUDA, "kernel cannot be resident (occupancy 0)");
    lp->block = block;
    lp->stride = staged ? stride : 0;
    lp->smem = (int)((staged ? smem : 0) + ksm);
    // split kernels: one trajectory per LANE of a 4-warp CTA
    const long long per_block = m->split ? 32 : (long long)block;
    const long long want = (N + per_block - 1) / per_block;
    lp->grid = (int)std::max<long long>(1, std::min<long long>((long long)nb * d->sms, want));
    int refill = o->refill_threshold;
    const bool lane_refill = (o->adaptive && !is_sde(m->alg)) || (m->flags & B200ENS_MODEL_SDE_ADAPTIVE);   // step counts vary per trajectory
    if (refill <= 0) refill = lane_refill ? 4 : 32;
    refill = std::min(refill, 32);
    lp->refill = refill;
    return 0;
}

int fill_args(const b200ens_model* m, const b200ens_opts* o, B2Args* a) {
    const int order = alg_order(m->alg);
    a->maxiters = o->maxiters > 0 ? o->maxiters : 100000;
    a->seed = o->seed;
    a->t0 = o->t0;
    a->t1 = o->t1;
    a->dt = o->dt;
    a->abstol = dflt(o->abstol, 1e-6);
    a->reltol = dflt(o->reltol, 1e-3);
    a->dtmin = dflt(o->dtmin, 0.0);
    a->dtmax = dflt(o->dtmax, o->t1 - o->t0);
    a->qmin = dflt(o->qmin, 0.2);
    const bool sde_adapt = (m->flags & B200ENS_MODEL_SDE_ADAPTIVE) != 0;   // strong order 3/2, StochasticDiffEq's qmax
    a->qmax = dflt(o->qmax, sde_adapt ? 1.125 : 10.0);
    a->gamma = dflt(o->gamma, 0.9);
    a->beta1 = dflt(o->beta1, sde_adapt ? 7.0 / 15.0 : 7.0 / (10.0 * order));
    a->beta2 = dflt(o->beta2, sde_adapt ? 4.0 / 15.0 : 2.0 / (5.0 * order));
    a->qoldinit = dflt(o->qoldinit, 1e-4);
    a->adaptive = is_sde(m->alg) ? 0 : (o->adaptive != 0);
    a->noise_injected = o->noise_injected;
    a->event_terminate = o->event_terminate;
    a->interp_points = o->interp_points > 0 ? o->interp_points : 10;
    int st = o->save_tstops;
    // auto: interpolate (every stepper has a dense output).  FBDF on a mass-matrix problem saves at tstops: its Hermite dense
    // output takes f for u', which M u' = f does not give for algebraic components (the Rodas dense output works on the stage
    // increments and is used for DAEs too)
    if (st < 0) st = (m->has_mass && m->alg == B200ENS_FBDF) ? 1 : 0;
    a->save_tstops = st;
    if (!(o->t1 > o->t0)) return fail(B200ENS_E_INVALID, "tspan must satisfy t1 > t0 (forward integration only)");
    if (!(o->dt > 0) && !(a->adaptive && o->dt == 0))
        return fail(B200ENS_E_INVALID, "dt must be > 0 (fixed step, SDE) or 0 = automatic initial step (adaptive)");
    a->f_t0 = (float)a->t0;
    a->f_t1 = (float)a->t1;
    a->f_dt = (float)a->dt;
    a->f_abstol = (float)a->abstol;
    a->f_reltol = (float)a->reltol;
    a->f_dtmin = (float)a->dtmin;
    a->f_dtmax = (float)a->dtmax;
    a->f_qmin = (float)a->qmin;
    a->f_qmax = (float)a->qmax;
    a->f_gamma = (float)a->gamma;
    a->f_beta1 = (float)a->beta1;
    a->f_beta2 = (float)a->beta2;
    a->f_qoldinit = (float)a->qoldinit;
    for (int i = 0; i < 32; i++) {
        const int c = i < m->n_state ? i : m->n_state - 1;
        const double ta = o->abstol_vec ? o->abstol_vec[c] : a->abstol, tr = o->reltol_vec ? o->reltol_vec[c] : a->reltol;
        if (!(ta >= 0) || !(tr >= 0) || !(ta + tr > 0)) return fail(B200ENS_E_INVALID, "tolerances must be >= 0 and not both 0 (component %d)", c);
        a->tol_a[i] = ta;
        a->tol_r[i] = tr;
        a->f_tol_a[i] = (float)ta;
        a->f_tol_r[i] = (float)tr;
    }
    a->nsteps_noise = is_sde(m->alg) ? (long long)std::ceil((o->t1 - o->t0) / o->dt - 1e-9) : 0;
    if (is_sde(m->alg) && !sde_adapt && a->nsteps_noise >= (1ll << 31))
        return fail(B200ENS_E_INVALID, "fixed-step SDE solve with %lld steps (the step index is 32-bit)", a->nsteps_noise);
    if (sde_adapt) a->nsteps_noise = o->noise_injected ? o->noise_stream_len : 0;   // length of the injected stream of normals
    return 0;
}

int check_opts(const b200ens_model* m, const b200ens_opts* o, int n_save, const void* dW) {
    if (!m || !o) return fail(B200ENS_E_INVALID, "null model or opts");
    if (o->struct_size != sizeof(b200ens_opts))
        return fail(B200ENS_E_INVALID, "b200ens_opts.struct_size %u != %zu (ABI mismatch)", o->struct_size,
                    sizeof(b200ens_opts));
    if (n_save < 0) return fail(B200ENS_E_INVALID, "n_save < 0");
    if (o->save_everystep) {
        if (is_sde(m->alg) || m->split)
            return fail(B200ENS_E_UNSUPPORTED, "save_everystep needs an ODE stepper and the one-thread kernel (compile with "
                                               "B200ENS_MODEL_NOSPLIT for large systems; SDE: use a saveat grid)");
        if (n_save < 2) return fail(B200ENS_E_INVALID, "save_everystep: n_save is the capacity per trajectory and must be >= 2");
        if (o->stage_outputs > 0) return fail(B200ENS_E_UNSUPPORTED, "save_everystep with stage_outputs=1");
    }
    if (o->n_tstops < 0 || (o->n_tstops > 0 && !o->tstops)) return fail(B200ENS_E_INVALID, "n_tstops = %d with tstops = %p", o->n_tstops, (const void*)o->tstops);
    if (o->n_tstops > 0) {
        if (is_sde(m->alg)) return fail(B200ENS_E_UNSUPPORTED, "tstops with an SDE stepper (fixed-step solves hit their dt grid; adaptive SDE: not implemented)");
        for (int i = 0; i < o->n_tstops; i++) {
            if (o->tstops[i] != o->tstops[i]) return fail(B200ENS_E_INVALID, "tstops[%d] is NaN", i);
            if (i && !(o->tstops[i] > o->tstops[i - 1])) return fail(B200ENS_E_INVALID, "tstops must be strictly ascending (entry %d)", i);
        }
    }
    if (o->noise_injected && !dW) return fail(B200ENS_E_INVALID, "noise_injected=1 but dW is NULL");
    if ((m->flags & B200ENS_MODEL_SDE_ADAPTIVE) && o->stage_outputs > 0)
        return fail(B200ENS_E_UNSUPPORTED, "adaptive SDE kernels store their outputs directly (stage_outputs=1 is not available)");
    if (o->noise_injected && (m->flags & B200ENS_MODEL_SDE_ADAPTIVE) && o->noise_stream_len <= 0)
        return fail(B200ENS_E_INVALID, "adaptive SDE stepping with noise_injected=1 needs opts.noise_stream_len > 0 (dW = [N][noise_stream_len] standard normals)");
    return 0;
}

// solve(...; tstops): the entries inside (t0, t1), converted to the state type (what the kernels compare t against)
std::vector<char> tstops_bytes(const b200ens_model* m, const b200ens_opts* o, int* n_out) {
    std::vector<char> b;
    int k = 0;
    for (int i = 0; i < o->n_tstops; i++) {
        const double v = o->tstops[i];
        if (!(v > o->t0 && v < o->t1)) continue;
        if (m->dtype == B200ENS_F64) {
            b.insert(b.end(), (const char*)&v, (const char*)&v + 8);
        } else {
            const float f = (float)v;
            if (k && memcmp(&f, b.data() + (size_t)(k - 1) * 4, 4) == 0) continue;   // two doubles that round to one float
            b.insert(b.end(), (const char*)&f, (const char*)&f + 4);
        }
        k++;
    }
    *n_out = k;
    return b;
}

int launch(b200ens_model* m, const LaunchPlan& lp, const B2Args& a, cudaStream_t stream) {
    void* params[] = {(void*)&a};
    // the specialised entry keeps 32-bit output offsets in its Float32 save queue: fall back to the generic entry beyond 2^32 elements
    const bool off32_ok = m->dtype == B200ENS_F64 || (unsigned long long)a.N * (unsigned long long)a.n_save * m->n_out < (1ull << 32);
    const bool tstops_ok = !a.save_tstops || is_rosenbrock(m->alg);   // the Rosenbrock entry keeps save_tstops a run-time flag
    const char* fg = getenv("B200ENS_GENERIC_ENTRY");   // tests: the generic entry must give the specialised one's bits
    const bool force_generic = fg && atoi(fg) != 0;
    cudaKernel_t k = (!force_generic && m->kernel_adaptive && a.adaptive && tstops_ok && a.dt > 0 && a.stage_stride == 0 && off32_ok && !a.save_every && !a.mom_sum && a.n_tstops == 0) ? m->kernel_adaptive : m->kernel;   // fused moments, tstops: generic entry
    if (k != m->kernel && lp.smem > 48 * 1024)
        CU(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, lp.smem));
    CU(cudaLaunchKernel((const void*)k, dim3(lp.grid), dim3(lp.block), params, lp.smem, stream));
    return 0;
}

// ---------------------------------------------------------------- expected-work ordering (kernels/b2_work.cuh)
constexpr int kWorkBuckets = 1024, kWorkTile = 4;
// The order is established inside windows of consecutive trajectories (kernels/b2_work.cuh: keeps the output rows that
// are being written at any moment within L2's reach, so fewer sectors are evicted half-written).  Measured, 1M Lorenz
// trajectories (profiles/README.md): every window boundary costs lane coherence (64k windows: +9 % time), the DRAM
// traffic falls from 2.1x to 1.3-1.45x the algorithmic bytes; a window whose output footprint is ~70 MB is as fast as
// the global sort or slightly faster (f32 512k: 0.795 vs 0.804 ms, f64 256k: 1.426 vs 1.446 ms).  Rows that are whole
// 32-byte sectors (n_state * sizeof(T) a multiple of 32, e.g. the 16-species network) are never half-written: one
// global sort (windows cost config 5 10 %).  B200ENS_WORK_WINDOW overrides (0 = one global sort).
long long work_window(long long N, size_t out_bytes_per_traj, size_t row_bytes, int requested) {
    const long long tile = (long long)kWorkBuckets * kWorkTile;
    long long w = 1 << 18;
    static const char* e = getenv("B200ENS_WORK_WINDOW");
    if (requested > 1) {   // opts.work_order > 1: the caller's window
        w = requested;
    } else if (e) {
        w = atoll(e);
        if (w <= 0) w = N;
    } else if (row_bytes % 32 == 0 || out_bytes_per_traj == 0) {
        w = N;
    } else {
        while (w < (1ll << 24) && (size_t)(2 * w) * out_bytes_per_traj <= (72u << 20)) w *= 2;
    }
    w = std::max(tile, (w + tile - 1) / tile * tile);
    return w;
}
long long work_windows(long long N, long long window) { return std::max<long long>(1, (N + window - 1) / window); }
size_t work_head_bytes(long long nwin) { return 16 + 2 * (size_t)nwin * kWorkBuckets * sizeof(unsigned); }   // counter | histograms | cursors
size_t work_scratch_bytes(long long N) {   // sized for the smallest window (most histograms)
    const long long tile = (long long)kWorkBuckets * kWorkTile;
    return work_head_bytes(work_windows(N, tile)) + (size_t)N * (sizeof(unsigned) + sizeof(unsigned short)) + 16;
}
bool want_work_order(const b200ens_model* m, const b200ens_opts* o, const B2Args& a, long long N) {
    if (!m->k_work_keys || !a.adaptive || N >= (1ll << 32)) return false;
    if (const char* e = getenv("B200ENS_WORK_ORDER")) return atoi(e) != 0;   // experiments
    return o->work_order > 0 || (o->work_order < 0 && N >= 32768);
}
// Enqueues memset + b2_work_keys + b2_work_scatter on `stream`; points a->perm / a->work_counter into `scratch`.
// a->u0, a->p, a->N and the tolerances must be final.
int enqueue_work_order(b200ens_model* m, const b200ens_opts* o, DeviceCtx* d, B2Args* a, void* scratch, cudaStream_t stream) {
    const long long window = work_window(a->N, (size_t)a->n_save * m->n_out * m->elem(), (size_t)m->n_out * m->elem(), o->work_order);
    const long long nwin = work_windows(a->N, window);
    const size_t head = work_head_bytes(nwin);
    char* base = (char*)scratch;
    unsigned* hist = (unsigned*)(base + 16);
    unsigned* cursor = hist + nwin * kWorkBuckets;
    unsigned* perm = cursor + nwin * kWorkBuckets;
    unsigned short* keys = (unsigned short*)(perm + a->N);
    CU(cudaMemsetAsync(base, 0, head, stream));
    a->work_counter = (unsigned long long*)base;
    a->perm = nullptr;
    {
        long long win = window;
        void* params[] = {(void*)a, (void*)&keys, (void*)&hist, (void*)&win};
        // every block ends with one global atomic per non-empty bucket: fewer, longer-running blocks mean fewer atomics on
        // the ~100 hot histogram addresses (B200ENS_KEYS_GRID_MULT: blocks per SM, experiments)
        static const int mult = getenv("B200ENS_KEYS_GRID_MULT") ? std::max(1, atoi(getenv("B200ENS_KEYS_GRID_MULT"))) : 4;   // measured 1M Lorenz f32: 0.954 (8), 0.949 (4), 0.952 (2) ms per step
        const int grid = (int)std::max<long long>(1, std::min<long long>((a->N + 255) / 256, (long long)d->sms * mult));
        CU(cudaLaunchKernel((const void*)m->k_work_keys, dim3(grid), dim3(256), params, 0, stream));
    }
    {
        long long N = a->N, win = window;
        const unsigned short* ck = keys;
        const unsigned* ch = hist;
        void* params[] = {(void*)&N, (void*)&ck, (void*)&ch, (void*)&cursor, (void*)&perm, (void*)&win};
        const long long tile = (long long)kWorkBuckets * kWorkTile;
        CU(cudaLaunchKernel((const void*)m->k_work_scatter, dim3((unsigned)((N + tile - 1) / tile)), dim3(kWorkBuckets), params, 0,
                            stream));
    }
    a->perm = perm;
    return 0;
}

// ---------------------------------------------------------------- on-device ensemble moments (b2_moments.cuh)
struct MomentsKernel {
    std::mutex mu;
    std::vector<char> cubin;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel = nullptr;
};
MomentsKernel g_moments[2];

int moments_kernel(int f64, cudaKernel_t* out) {
    MomentsKernel& mk = g_moments[f64 ? 1 : 0];
    std::lock_guard<std::mutex> lk(mk.mu);
    if (!mk.kernel) {
        std::string src = std::string("#define B2M_F64 ") + (f64 ? "1" : "0") + "\n#include \"b2_moments.cuh\"\n";
        B2_NVRTC_OR_FAIL(nv);
        nvrtcProgram prog;
        const char* hdr_names[kNumHeaders];
        const char* hdr_text[kNumHeaders];
        for (int i = 0; i < kNumHeaders; i++) {
            hdr_names[i] = kHeaders[i].name;
            hdr_text[i] = kHeaders[i].text;
        }
        if (nv->CreateProgram(&prog, src.c_str(), "b200ens_moments.cu", kNumHeaders, hdr_text, hdr_names) != NVRTC_SUCCESS)
            return fail(B200ENS_E_COMPILE, "nv->CreateProgram(moments) failed");
        const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
        nvrtcResult r = nv->CompileProgram(prog, 3, opts);
        if (r != NVRTC_SUCCESS) {
            size_t n = 0;
            nv->GetProgramLogSize(prog, &n);
            std::string log(n, 0);
            if (n) nv->GetProgramLog(prog, &log[0]);
            nv->DestroyProgram(&prog);
            return fail(B200ENS_E_COMPILE, "NVRTC (moments kernel): %.1000s", log.c_str());
        }
        size_t n = 0;
        nv->GetCUBINSize(prog, &n);
        mk.cubin.resize(n);
        nv->GetCUBIN(prog, mk.cubin.data());
        nv->DestroyProgram(&prog);
        CU(cudaLibraryLoadData(&mk.lib, mk.cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
        CU(cudaLibraryGetKernel(&mk.kernel, mk.lib, "b2_moments_kernel"));
    }
    *out = mk.kernel;
    return 0;
}

struct Moments {       // per-shard request: accumulate instead of copying out_u back
    double* sum = nullptr;       // host [row_len], this shard's partial result
    double* sumsq = nullptr;
    long long count = 0;
};

struct ShardResult {
    int code = 0;
    std::string err;
    double h2d = 0, kern = 0, d2h = 0, total = 0;
    int launches = 0;
    int fused = 0, fused_fallbacks = 0;   // ensemble moments accumulated inside the solve kernel; chunks recomputed through out_u
    LaunchPlan lp;
};

// Solve the trajectory ranges [lo, hi) in `ranges` (the blocks dealt to this device) of the caller's host buffers on one device.
typedef std::vector<std::pair<long long, long long>> Ranges;
int solve_shard(b200ens_model* m, const b200ens_opts* o, int dev, const Ranges& ranges, const char* u0,
                const char* p, const void* saveat, int n_save, const char* dW, char* out_u, int32_t* retcode,
                b200ens_stats* stats, ShardResult* res, Moments* mom = nullptr, char* out_t_every = nullptr) {
    DeviceCtx* d;
    int rc = device_ctx(dev, &d);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(d->mu);
    CU(cudaSetDevice(dev));
    rc = ensure_loaded(m);
    if (rc) return rc;
    const size_t es = m->elem();
    const int n = m->n_state, np = m->n_param;
    B2Args base{};
    rc = fill_args(m, o, &base);
    if (rc) return rc;
    const bool sde_adapt = (m->flags & B200ENS_MODEL_SDE_ADAPTIVE) != 0;   // dW = [N][noise_stream_len] standard normals
    const size_t noise_per_traj = !dW ? 0 : sde_adapt ? (size_t)base.nsteps_noise * es
                                  : (size_t)base.nsteps_noise * ((m->alg == B200ENS_SOSRA || m->alg == B200ENS_SRIW1) ? 2 : 1) * n * es;
    const size_t out_per_traj = (size_t)n_save * m->n_out * es;

    rc = grow(&d->saveat, &d->cap_save, std::max<size_t>(es, (size_t)n_save * es));
    if (rc) return rc;
    const auto wall0 = std::chrono::steady_clock::now();
    // B200ENS_TRACE=1: host-side timeline of the pipeline (ms since entry) on stderr -- where the wall time of an
    // end-to-end solve goes beyond the PCIe floor
    static const bool trace = getenv("B200ENS_TRACE") != nullptr;
    std::string trace_txt;
    auto mark = [&](const char* what, long long v = -1) {
        if (!trace) return;
        char b[96];
        snprintf(b, sizeof b, " %s%s%s@%.3f", what, v >= 0 ? ":" : "", v >= 0 ? std::to_string(v).c_str() : "",
                 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count());
        trace_txt += b;
    };
    const bool every = o->save_everystep != 0;   // n_save = capacity, no saveat grid
    if (n_save && !every) CU(cudaMemcpyAsync(d->saveat, saveat, (size_t)n_save * es, cudaMemcpyHostToDevice, d->slot[0].stream));
    int n_ts = 0;
    const std::vector<char> ts_host = tstops_bytes(m, o, &n_ts);
    if (n_ts) {
        if ((rc = grow(&d->tstops, &d->cap_tstops, ts_host.size()))) return rc;
        CU(cudaMemcpyAsync(d->tstops, ts_host.data(), ts_host.size(), cudaMemcpyHostToDevice, d->slot[0].stream));
    }
    base.tstops = n_ts ? d->tstops : nullptr;
    base.n_tstops = n_ts;
    CU(cudaStreamSynchronize(d->slot[0].stream));

    // chunk size: keep both pipeline slots under ~1/4 of the device memory and at least a few waves
    long long total = 0;
    for (const auto& r : ranges) total += r.second - r.first;
    const size_t per_traj = (size_t)n * es + (size_t)np * es + out_per_traj + noise_per_traj + 4 + sizeof(b200ens_stats);
    size_t free_b = 0, total_b = 0;
    mark("saveat");
    // cudaMemGetInfo costs ~0.09 ms (a tenth of the kernel time of 1M Lorenz trajectories): only ask when four slots of
    // the largest chunk could come anywhere near the device memory (total size cached at device init)
    if (4 * per_traj * (size_t)std::min<long long>(total, 1 << 20) > d->mem_total / 8) {
        CU(cudaMemGetInfo(&free_b, &total_b));
    } else {
        free_b = total_b = d->mem_total;
    }
    mark("meminfo");
    // Four pipeline slots (stream + buffers each); everything is enqueued without host synchronisation.  The run is
    // D2H-bound (1M Float32 Lorenz trajectories: 152 MB back = 2.7 ms of PCIe against 1.0 ms of kernel), so the
    // schedule starts SMALL and grows: 1/16 of the shard first (the D2H engine has work after ~0.3 ms instead of
    // ~0.6 ms with equal quarters), then 3/16, then the rest in equal large chunks (large copies reach 53-56 GB/s;
    // per-chunk host overhead is ~30 us of API calls; B200ENS_TRACE=1 prints the timeline).
    long long chunk = std::min<long long>(total, 1 << 20);
    const long long mem_cap = (long long)((free_b / 8) / std::max<size_t>(1, per_traj));
    chunk = std::max<long long>(1, std::min(chunk, mem_cap));
    // A chunk = one kernel launch over `cn` trajectories that sit contiguously in the DEVICE buffers; on the host it is
    // one or more pieces (the blocks dealt to this device, multi-device solves): every piece is H2D-copied to its offset
    // in the chunk and D2H-copied straight back to its own place in the caller's arrays.  So a device launches the same
    // few large kernels whether its share of the ensemble is one range or many interleaved blocks.
    struct Piece {
        long long g0, cn, off;   // first global trajectory, count, offset inside the chunk
    };
    struct Chunk {
        long long cn = 0;
        std::vector<Piece> pieces;
    };
    std::vector<long long> sched;
    if (const char* e = getenv("B200ENS_CHUNK")) {   // fixed size (tests, experiments)
        chunk = std::max<long long>(1, atoll(e));
        for (long long r = total; r > 0; r -= chunk) sched.push_back(std::min(chunk, r));
    } else {
        long long frac = 16;
        if (const char* e = getenv("B200ENS_FIRST_FRAC")) frac = std::max(2, atoi(e));   // experiments
        const long long first = std::min<long long>(chunk, std::max<long long>(1 << 14, (total + frac - 1) / frac));
        long long r = total;
        if (r > 4 * first) {
            sched.push_back(first);
            r -= first;
            const long long second = std::min<long long>(chunk, 3 * first);
            sched.push_back(second);
            r -= second;
        }
        const long long k = std::max<long long>(r > 2 * first ? 2 : 1, (r + chunk - 1) / chunk);
        for (long long i = 0; i < k; i++) {
            const long long c = (r + (k - i) - 1) / (k - i);
            sched.push_back(c);
            r -= c;
        }
        chunk = *std::max_element(sched.begin(), sched.end());
    }
    std::vector<Chunk> work;   // in launch order; the chunk sizes follow `sched` over the concatenation of the ranges
    {
        size_t ri = 0;
        long long rpos = ranges.empty() ? 0 : ranges[0].first;
        long long left = total;
        for (size_t i = 0; left > 0; i++) {
            Chunk c;
            long long want = std::min<long long>(sched[std::min(i, sched.size() - 1)], left);
            while (want > 0) {
                const long long take = std::min(want, ranges[ri].second - rpos);
                c.pieces.push_back({rpos, take, c.cn});
                c.cn += take;
                want -= take;
                rpos += take;
                if (rpos == ranges[ri].second && ++ri < ranges.size()) rpos = ranges[ri].first;
            }
            left -= c.cn;
            work.push_back(std::move(c));
        }
    }

    LaunchPlan lp;
    rc = plan_launch(m, o, d, chunk, n_save, &lp);
    if (rc) return rc;
    res->lp = lp;
    // moments mode: device accumulators [sum | sumsq | count] instead of the D2H copy of out_u
    const int row_len = n_save * m->n_out;
    cudaKernel_t mom_kernel = nullptr;
    double* d_acc = nullptr;
    // FUSED mode: the ODE kernels add every saved value to the accumulators themselves (global double reductions) and
    // out_u is never allocated, written or read back.  Measured on B200 (profiles/README.md, round 2): it removes the
    // out_u traffic entirely (config 5 with 1001 save points: 2 x 12.8 GB per 100k trajectories -> 0.5 MB) but the L2
    // sustains ~1e11 Float64 reductions per second when they are spread over thousands of addresses and 3e10 on a few
    // thousand, so the fused solve is 18 % SLOWER than solve + second pass (b2_moments.cuh) on config 5 and 5x slower
    // on 1M Lorenz trajectories x 401 save points: HBM absorbs the rows faster than L2 adds them.  Hence the second
    // pass is the default; the fused mode is taken when a trajectory's rows are so large (>= 4 MB) that out_u would
    // force tiny chunks, or on request (B200ENS_FUSE_MOMENTS=1; =0 forbids it).  Only successful trajectories may
    // count: a chunk in which any trajectory failed is recomputed through out_u (rare; mom_fail tells).
    bool fuse = mom && !is_sde(m->alg) && !every && out_per_traj >= ((size_t)4 << 20);
    if (const char* e = getenv("B200ENS_FUSE_MOMENTS")) fuse = mom && !is_sde(m->alg) && !every && row_len > 0 && atoi(e) != 0;
    const size_t acc_bytes = (2 * (size_t)row_len + 1) * sizeof(double);
    if (mom) {
        if ((rc = moments_kernel(m->dtype == B200ENS_F64, &mom_kernel))) return rc;
        if ((rc = grow(&d->acc, &d->cap_acc, acc_bytes))) return rc;
        d_acc = static_cast<double*>(d->acc);
        CU(cudaMemsetAsync(d_acc, 0, acc_bytes, d->slot[0].stream));
        CU(cudaStreamSynchronize(d->slot[0].stream));  // both pipeline streams accumulate into it
        for (int i = 0; i < row_len; i++) mom->sum[i] = mom->sumsq[i] = 0.0;
        mom->count = 0;
    }

    // Callers with pageable arrays (plain Julia Arrays, numpy): stage through pinned bounce buffers with a parallel
    // host memcpy (cudaMemcpyAsync straight from pageable memory measured 10x slower, profiles/r1_e2e_probe.log).
    const bool stage_in = !(is_pinned(u0) && is_pinned(p) && is_pinned(dW));
    const bool stage_out = !(is_pinned(mom ? nullptr : out_u) && is_pinned(retcode) && is_pinned(stats));
    struct Pending {
        bool used = false;
        const Chunk* ck = nullptr;
        B2Args args{};     // fused moments: what the chunk was launched with (the fallback relaunches it through out_u)
        LaunchPlan lp;
    } pend[kMaxSlots];
    int nslots = kMaxSlots;
    if (const char* e = getenv("B200ENS_SLOTS")) nslots = std::max(1, std::min(kMaxSlots, atoi(e)));
    int it = 0;
    const size_t out_b = (mom || !n_save) ? 0 : out_per_traj;   // bytes per trajectory in the staged output block
    auto launch_second_pass = [&](Slot& s, long long cn, double* acc) -> int {   // b2_moments.cuh over s.out / s.rc
        long long nn = cn;
        int rl = row_len;
        const void* po = s.out;
        const int* prc = s.rc;
        double* ps = acc;
        double* pq = acc + row_len;
        unsigned long long* pc = reinterpret_cast<unsigned long long*>(acc + 2 * (size_t)row_len);
        void* margs[] = {&po, &prc, &nn, &rl, &ps, &pq, &pc};
        const int gx = (row_len + 127) / 128;
        const int gy = (int)std::max<long long>(1, std::min<long long>(cn, (long long)d->sms * 16 / gx));
        CU(cudaLaunchKernel((const void*)mom_kernel, dim3(gx, gy), dim3(128), margs, 0, s.stream));
        res->launches++;
        return 0;
    };
    auto collect = [&](Slot& s, const Pending& pd) -> int {
        CU(cudaEventSynchronize(s.ev[3]));
        if (fuse) {
            const double* h = reinterpret_cast<const double*>(s.h_macc);
            unsigned long long fails;
            memcpy(&fails, h + 2 * (size_t)row_len, sizeof fails);
            unsigned long long cnt = (unsigned long long)pd.ck->cn;
            if (fails) {
                // some trajectory of this chunk failed after contributing: recompute the chunk through out_u and the
                // second pass (inputs, permutation and retcode buffers of the slot are still in place)
                int r2;
                if ((r2 = grow(&s.out, &s.cap_out, std::max<size_t>(es, (size_t)pd.ck->cn * out_per_traj)))) return r2;
                B2Args a2 = pd.args;
                a2.out_u = s.out;
                a2.mom_sum = a2.mom_sq = nullptr;
                a2.mom_fail = nullptr;
                CU(cudaMemsetAsync(a2.work_counter, 0, sizeof(unsigned long long), s.stream));
                if ((r2 = launch(m, pd.lp, a2, s.stream))) return r2;
                res->launches++;
                CU(cudaMemsetAsync(s.macc, 0, acc_bytes, s.stream));
                if ((r2 = launch_second_pass(s, pd.ck->cn, static_cast<double*>(s.macc)))) return r2;
                CU(cudaMemcpyAsync(s.h_macc, s.macc, acc_bytes, cudaMemcpyDeviceToHost, s.stream));
                CU(cudaStreamSynchronize(s.stream));
                memcpy(&cnt, h + 2 * (size_t)row_len, sizeof cnt);
                res->fused_fallbacks++;
            }
            for (int i = 0; i < row_len; i++) {
                mom->sum[i] += h[i];
                mom->sumsq[i] += h[row_len + i];
            }
            mom->count += (long long)cnt;
        }
        if (stage_out) {
            const long long ccn = pd.ck->cn;
            const char* h_o = s.h_out;
            const char* h_rc = h_o + (size_t)ccn * out_b;
            const char* h_st = h_rc + (size_t)ccn * sizeof(int);
            for (const Piece& pc : pd.ck->pieces) {
                if (out_b) par_memcpy(out_u + (size_t)pc.g0 * out_per_traj, h_o + (size_t)pc.off * out_b, (size_t)pc.cn * out_b);
                memcpy(retcode + pc.g0, h_rc + (size_t)pc.off * sizeof(int), (size_t)pc.cn * sizeof(int));
                if (stats) par_memcpy(stats + pc.g0, h_st + (size_t)pc.off * sizeof(b200ens_stats), (size_t)pc.cn * sizeof(b200ens_stats));
            }
        }
        float a_ = 0, b_ = 0, c_ = 0;
        CU(cudaEventElapsedTime(&a_, s.ev[0], s.ev[1]));
        CU(cudaEventElapsedTime(&b_, s.ev[1], s.ev[2]));
        CU(cudaEventElapsedTime(&c_, s.ev[2], s.ev[3]));
        res->h2d += a_;
        res->kern += b_;
        res->d2h += c_;
        return 0;
    };
    for (const Chunk& wk : work) {
        const long long cn = wk.cn;
        const long long g0 = wk.pieces[0].g0;  // global index of the chunk's first trajectory (Philox streams: SDE solves are
                                               // dealt as ONE contiguous range per device, so the chunk is contiguous too)
        Slot& s = d->slot[it % nslots];
        if (pend[it % nslots].used) {
            rc = collect(s, pend[it % nslots]);
            if (rc) return rc;
        }
        if ((rc = grow(&s.u0, &s.cap_u0, (size_t)cn * n * es))) return rc;
        if ((rc = grow(&s.p, &s.cap_p, std::max<size_t>(es, (size_t)cn * np * es)))) return rc;
        if (!fuse && (rc = grow(&s.out, &s.cap_out, std::max<size_t>(es, (size_t)cn * out_per_traj)))) return rc;
        if (fuse) {
            if ((rc = grow(&s.macc, &s.cap_macc, acc_bytes))) return rc;
            if ((rc = grow_host(&s.h_macc, &s.cap_hmacc, acc_bytes))) return rc;
        }
        if (dW && (rc = grow(&s.dW, &s.cap_dW, (size_t)cn * noise_per_traj))) return rc;
        if (every && (rc = grow(&s.every_t, &s.cap_every, (size_t)cn * n_save * es))) return rc;
        if ((size_t)cn > s.cap_n) {
            if (s.rc) CU(cudaFree(s.rc));
            if (s.stats) CU(cudaFree(s.stats));
            s.rc = nullptr;
            s.stats = nullptr;
            CU(cudaMalloc(&s.rc, (size_t)cn * sizeof(int)));
            CU(cudaMalloc(&s.stats, (size_t)cn * sizeof(b200ens_stats)));
            s.cap_n = (size_t)cn;
        }
        if (stage_in) {
            const size_t bu = (size_t)cn * n * es, bp = (size_t)cn * np * es, bw = dW ? (size_t)cn * noise_per_traj : 0;
            if ((rc = grow_host(&s.h_in, &s.cap_hin, bu + bp + bw))) return rc;
            for (const Piece& pc : wk.pieces) {
                par_memcpy(s.h_in + (size_t)pc.off * n * es, u0 + (size_t)pc.g0 * n * es, (size_t)pc.cn * n * es);
                if (bp) par_memcpy(s.h_in + bu + (size_t)pc.off * np * es, p + (size_t)pc.g0 * np * es, (size_t)pc.cn * np * es);
                if (bw) par_memcpy(s.h_in + bu + bp + (size_t)pc.off * noise_per_traj, dW + (size_t)pc.g0 * noise_per_traj, (size_t)pc.cn * noise_per_traj);
            }
        }
        if (stage_out && (rc = grow_host(&s.h_out, &s.cap_hout, (size_t)cn * (out_b + sizeof(int) + (stats ? sizeof(b200ens_stats) : 0)))))
            return rc;
        CU(cudaEventRecord(s.ev[0], s.stream));
        if (stage_in) {
            const size_t bu = (size_t)cn * n * es, bp = (size_t)cn * np * es;
            CU(cudaMemcpyAsync(s.u0, s.h_in, bu, cudaMemcpyHostToDevice, s.stream));
            if (np) CU(cudaMemcpyAsync(s.p, s.h_in + bu, bp, cudaMemcpyHostToDevice, s.stream));
            if (dW) CU(cudaMemcpyAsync(s.dW, s.h_in + bu + bp, (size_t)cn * noise_per_traj, cudaMemcpyHostToDevice, s.stream));
        } else {
            for (const Piece& pc : wk.pieces) {
                CU(cudaMemcpyAsync((char*)s.u0 + (size_t)pc.off * n * es, u0 + (size_t)pc.g0 * n * es, (size_t)pc.cn * n * es, cudaMemcpyHostToDevice, s.stream));
                if (np) CU(cudaMemcpyAsync((char*)s.p + (size_t)pc.off * np * es, p + (size_t)pc.g0 * np * es, (size_t)pc.cn * np * es, cudaMemcpyHostToDevice, s.stream));
                if (dW) CU(cudaMemcpyAsync((char*)s.dW + (size_t)pc.off * noise_per_traj, dW + (size_t)pc.g0 * noise_per_traj, (size_t)pc.cn * noise_per_traj, cudaMemcpyHostToDevice, s.stream));
            }
        }
        B2Args a = base;
        a.u0 = s.u0;
        a.p = s.p;
        a.saveat = every ? nullptr : d->saveat;
        a.every_t = every ? s.every_t : nullptr;
        a.save_every = every ? 1 : 0;
        a.dW = dW ? s.dW : nullptr;
        a.out_u = s.out;
        a.retcode = s.rc;
        a.stats = stats ? s.stats : nullptr;
        a.work_counter = s.counter;
        a.N = cn;
        a.traj_offset = o->traj_offset + (unsigned long long)g0;
        a.n_save = n_save;
        a.refill_threshold = lp.refill;
        a.stage_stride = lp.stride;
        if (fuse) {
            a.out_u = nullptr;
            a.stage_stride = 0;   // nothing to stage or flush: the sink adds instead of storing
            a.mom_sum = static_cast<double*>(s.macc);
            a.mom_sq = a.mom_sum + row_len;
            a.mom_fail = reinterpret_cast<unsigned long long*>(a.mom_sum + 2 * (size_t)row_len);
            CU(cudaMemsetAsync(s.macc, 0, acc_bytes, s.stream));
        }
        LaunchPlan l2 = lp;
        const long long pb = m->split ? 32 : (long long)lp.block;
        l2.grid = (int)std::max<long long>(1, std::min<long long>(lp.grid, (cn + pb - 1) / pb));
        CU(cudaEventRecord(s.ev[1], s.stream));
        if (want_work_order(m, o, a, cn)) {
            if ((rc = grow(&s.work, &s.cap_work, work_scratch_bytes(cn)))) return rc;
            if ((rc = enqueue_work_order(m, o, d, &a, s.work, s.stream))) return rc;
            res->launches += 2;
        } else {
            CU(cudaMemsetAsync(s.counter, 0, sizeof(unsigned long long), s.stream));
        }
        if ((rc = launch(m, l2, a, s.stream))) return rc;
        CU(cudaEventRecord(s.ev[2], s.stream));
        if (fuse) {
            CU(cudaMemcpyAsync(s.h_macc, s.macc, acc_bytes, cudaMemcpyDeviceToHost, s.stream));
        } else if (mom) {
            if ((rc = launch_second_pass(s, cn, d_acc))) return rc;
        }
        if (stage_out) {
            char* dst_rc = s.h_out + (size_t)cn * out_b;
            char* dst_st = dst_rc + (size_t)cn * sizeof(int);
            if (out_b) CU(cudaMemcpyAsync(s.h_out, s.out, (size_t)cn * out_per_traj, cudaMemcpyDeviceToHost, s.stream));
            CU(cudaMemcpyAsync(dst_rc, s.rc, (size_t)cn * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            if (stats) CU(cudaMemcpyAsync(dst_st, s.stats, (size_t)cn * sizeof(b200ens_stats), cudaMemcpyDeviceToHost, s.stream));
        } else {
            for (const Piece& pc : wk.pieces) {
                if (out_b) CU(cudaMemcpyAsync(out_u + (size_t)pc.g0 * out_per_traj, (char*)s.out + (size_t)pc.off * out_per_traj, (size_t)pc.cn * out_per_traj, cudaMemcpyDeviceToHost, s.stream));
                CU(cudaMemcpyAsync(retcode + pc.g0, s.rc + pc.off, (size_t)pc.cn * sizeof(int), cudaMemcpyDeviceToHost, s.stream));
                if (stats) CU(cudaMemcpyAsync(stats + pc.g0, s.stats + pc.off, (size_t)pc.cn * sizeof(b200ens_stats), cudaMemcpyDeviceToHost, s.stream));
            }
        }
        if (every)   // step times straight into the caller's [N][capacity] array
            for (const Piece& pc : wk.pieces)
                CU(cudaMemcpyAsync(out_t_every + (size_t)pc.g0 * n_save * es, (char*)s.every_t + (size_t)pc.off * n_save * es, (size_t)pc.cn * n_save * es, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.ev[3], s.stream));
        mark("enq", cn);
        pend[it % nslots].used = true;
        pend[it % nslots].ck = &wk;
        pend[it % nslots].args = a;
        pend[it % nslots].lp = l2;
        res->launches++;
        it++;
    }
    for (int k = 0; k < kMaxSlots; k++) {
        if (pend[k].used && (rc = collect(d->slot[k], pend[k]))) return rc;
        mark("collect", k);
    }
    if (trace) fprintf(stderr, "[b200ens trace dev %d]%s\n", dev, trace_txt.c_str());
    if (mom && !fuse) {
        std::vector<double> h(2 * (size_t)row_len + 1);
        CU(cudaMemcpy(h.data(), d_acc, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < row_len; i++) {
            mom->sum[i] = h[i];
            mom->sumsq[i] = h[row_len + i];
        }
        unsigned long long c;
        memcpy(&c, &h[2 * (size_t)row_len], sizeof c);
        mom->count = (long long)c;
    }
    res->fused = fuse ? 1 : 0;
    CU(cudaGetLastError());
    res->total = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    return 0;
}

}  // namespace

extern "C" {

int b200ens_abi_version(void) { return B200ENS_ABI_VERSION; }

int b200ens_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorStubLibrary) {
        (void)cudaGetLastError();
        return 0;
    }
    if (e != cudaSuccess) {
        fail(B200ENS_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        (void)cudaGetLastError();
        return -1;
    }
    return n;
}

const char* b200ens_last_error(void) { return g_err.c_str(); }

void b200ens_opts_init(b200ens_opts* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->struct_size = sizeof *o;
    o->adaptive = 1;
    o->abstol = o->reltol = o->dtmin = o->dtmax = -1;
    o->qmin = o->qmax = o->gamma = o->beta1 = o->beta2 = o->qoldinit = -1;
    o->maxiters = 0;
    o->interp_points = 0;
    o->save_tstops = -1;
    o->stage_outputs = -1;
    o->work_order = -1;
}

int b200ens_compile(const b200ens_model_desc* d, b200ens_model** out, char* log, size_t log_len) {
    if (log && log_len) log[0] = 0;
    if (!d || !out) return fail(B200ENS_E_INVALID, "null argument");
    *out = nullptr;
    if (d->struct_size != sizeof(b200ens_model_desc))
        return fail(B200ENS_E_INVALID, "b200ens_model_desc.struct_size %u != %zu (ABI mismatch)", d->struct_size,
                    sizeof(b200ens_model_desc));
    if (d->n_state < 1 || d->n_state > 32) return fail(B200ENS_E_INVALID, "n_state must be in 1..32");
    if (d->n_param < 0 || d->n_param > 64) return fail(B200ENS_E_INVALID, "n_param must be in 0..64");
    if (d->dtype != B200ENS_F32 && d->dtype != B200ENS_F64) return fail(B200ENS_E_INVALID, "bad dtype");
    if (d->alg < B200ENS_TSIT5 || d->alg > B200ENS_FBDF) return fail(B200ENS_E_INVALID, "bad alg id %d", d->alg);
    if (!d->rhs_src) return fail(B200ENS_E_INVALID, "rhs_src is required");
    if (needs_jac(d->alg) && !d->jac_src)
        return fail(B200ENS_E_UNSUPPORTED, "Rosenbrock methods and FBDF need the analytic Jacobian (jac_src); there is no AD/finite-difference fallback");
    if (is_sde(d->alg) && !d->noise_src) return fail(B200ENS_E_INVALID, "SDE algorithms need noise_src");
    if ((d->flags & B200ENS_MODEL_SDE_ADAPTIVE) && d->alg != B200ENS_SOSRA && d->alg != B200ENS_SRIW1)
        return fail(B200ENS_E_UNSUPPORTED, "B200ENS_MODEL_SDE_ADAPTIVE needs a stepper with an embedded error estimate: SRIW1 or SOSRA");
    if ((d->condition_src != nullptr) != (d->affect_src != nullptr))
        return fail(B200ENS_E_INVALID, "condition_src and affect_src must be given together");
    if ((d->dcondition_src != nullptr) != (d->daffect_src != nullptr))
        return fail(B200ENS_E_INVALID, "dcondition_src and daffect_src must be given together");
    if (is_sde(d->alg) && (d->condition_src || d->dcondition_src))
        return fail(B200ENS_E_UNSUPPORTED, "callbacks on SDE algorithms are not supported");
    if (d->n_save_idxs < 0 || d->n_save_idxs > d->n_state || (d->n_save_idxs > 0 && !d->save_idxs))
        return fail(B200ENS_E_INVALID, "n_save_idxs = %d (n_state %d)", d->n_save_idxs, d->n_state);
    for (int i = 0; i < d->n_save_idxs; i++) {
        if (d->save_idxs[i] < 0 || d->save_idxs[i] >= d->n_state) return fail(B200ENS_E_INVALID, "save_idxs[%d] = %d is not a component of the state", i, d->save_idxs[i]);
        for (int j = 0; j < i; j++)
            if (d->save_idxs[j] == d->save_idxs[i]) return fail(B200ENS_E_INVALID, "save_idxs holds component %d twice", d->save_idxs[i]);
    }
    auto m = std::make_unique<b200ens_model>();
    m->n_state = d->n_state;
    m->n_out = d->n_save_idxs > 0 ? d->n_save_idxs : d->n_state;
    m->n_param = d->n_param;
    m->dtype = d->dtype;
    m->alg = d->alg;
    m->has_mass = strstr(d->rhs_src, "#define B2_HAS_MASS 1") != nullptr;
    m->flags = d->flags;
    m->has_event = d->condition_src != nullptr;
    m->has_noise = d->noise_src != nullptr;
    m->name = d->name ? d->name : "";
    // Occupancy target for __launch_bounds__(128, min_blocks): start from the resident-CTA count that was
    // fastest on B200 for register-light models (profiles/: Tsit5 Lorenz f32 7 CTAs = 70 regs, f64 5 CTAs =
    // 96 regs) and back off until ptxas no longer spills more than a few words.
    int mb = d->dtype == B200ENS_F64 ? 5 : 7;
    if (const char* e = getenv("B200ENS_MINBLOCKS")) mb = std::max(1, atoi(e));
    int rc = 0;
    const char* force_k = getenv("B200ENS_KSMEM");
    const int nvec = d->alg == B200ENS_TSIT5 ? 7 : d->alg == B200ENS_VERN7 ? 14 : 0;
    const bool flag_k = (d->flags & B200ENS_MODEL_KSMEM) != 0;
    bool try_regs = !(((force_k && atoi(force_k) == 1) || flag_k) && nvec);
    // Vern7 on a system whose 14 stage vectors need more than ~200 registers cannot fit one thread: compile the split
    // kernel first and skip the (slow to compile, several hundred KB) one-thread variants when it builds within its spill
    // budget.  Same choice as the general rule below makes for these models, at half the compile time.
    bool split_tried = false;
    auto try_split = [&](int keep_spill, bool have_keep, bool forced) -> bool {
        int mbs = 3, rc2 = 0;
        if (const char* e = getenv("B200ENS_MINBLOCKS")) mbs = std::max(1, atoi(e));
        for (;; mbs--) {
            m->source = build_source(d, mbs, 128, 0, 1);
            rc2 = nvrtc_compile(m.get());
            if (rc2) break;
            parse_ptxas_log(m.get());
            if (m->spill <= 2048 || mbs <= 1 || getenv("B200ENS_MINBLOCKS")) break;
        }
        split_tried = true;
        if (!rc2 && (forced || (have_keep ? m->spill < keep_spill : m->spill <= 2048))) {
            m->split = 1;
            m->block = 128;
            mb = mbs;
            return true;
        }
        return false;
    };
    const char* force_s = getenv("B200ENS_SPLIT");
    const bool split_off = (d->flags & B200ENS_MODEL_NOSPLIT) || (force_s && atoi(force_s) == 0);
    const bool split_on = (d->flags & B200ENS_MODEL_SPLIT) || (force_s && atoi(force_s) == 1);
    const bool split_eligible = nvec && !d->dcondition_src && d->n_state >= 4 && !flag_k &&   // (Continuous and VectorContinuous callbacks: both kernels)
                                !(force_k && atoi(force_k) == 1);
    const bool surely_spills = d->alg == B200ENS_VERN7 && nvec * d->n_state * (d->dtype == B200ENS_F64 ? 2 : 1) > 200;   // + ~55 registers for everything else > 255
    if (try_regs && split_eligible && !split_off && (split_on || surely_spills) && try_split(0, false, split_on)) {
        try_regs = false;   // settled
        rc = 0;
    }
    // B200ENS_BLOCK: threads per CTA of the one-thread kernels (experiments; default 128)
    int blk = kBlock;
    // stiff steppers beyond 10 states keep W and the stage vectors in 2-3.5 KB of local memory per thread: 64 threads per SM
    // (one CTA of 64) keep that inside the L1 (Rodas5P, n = 16: 111 ms per 100k trajectories against 144 ms with 128 threads)
    // (the rolled-LU sizes only, n > 10: at n = 10 the unrolled kernel runs 2x faster with the usual 128 threads x 2 CTAs)
    if (needs_jac(d->alg) && d->n_state > 10) blk = 64;
    if (const char* e = getenv("B200ENS_BLOCK")) blk = std::max(32, std::min(256, atoi(e) / 32 * 32));
    if (try_regs) {
        m->block = blk;
        for (;; mb--) {
            m->source = build_source(d, mb, blk, 0);
            rc = nvrtc_compile(m.get());
            if (rc) break;
            parse_ptxas_log(m.get());
            if (m->spill <= 96 || mb <= 1 || getenv("B200ENS_MINBLOCKS")) break;
        }
    }
    // Large systems: the k-vectors do not fit the register file (config 5: n=16 f64 Vern7 spills 4 KB per thread).
    // When the register variant spills more than 4 KB (or on request: B200ENS_MODEL_KSMEM / B200ENS_KSMEM=1) compile
    // with the stage vectors in shared memory, CTA size chosen so that two CTAs fit in the 227 KB of an SM, and keep
    // it if it spills less.  Measured on config 5 (profiles/README.md): 156 vs 204 ms (saveat 101), 234 vs 297 ms
    // (saveat 1001) per 200k trajectories.
    // Large systems: SPLIT the trajectory over the four warps of a CTA (kernels/b2_split.cuh): a thread holds a quarter
    // of every state / stage vector, and the generated RHS exists once per warp role out of line, which keeps the loop
    // body inside the instruction cache.  Measured on the 16-species network, Float64, 200k trajectories
    // (profiles/README.md): Vern7 + event 42 ms split vs 103 ms one-thread (shared-memory stage vectors);
    // Vern7 without event 6.4 vs 7.8 ms; Tsit5 without event 4.4 vs 3.9 ms (its 7 stage vectors nearly fit one thread).
    // Tried when the one-thread variant spills more than 1 KB (Vern7; its loop body is then several hundred KB of SASS)
    // or 4 KB (Tsit5), kept when it spills less; B200ENS_MODEL_SPLIT / B200ENS_SPLIT=1 force it.
    // Unlike the one-thread kernels the split kernel prefers occupancy over a spill-free build (four warps meet at a
    // barrier ~25 times per step): start at 3 CTAs/SM and accept up to 2 KB of spill stores.
    if (!rc && !m->split && !split_tried && split_eligible && !split_off &&
        (split_on || (try_regs && m->spill > (d->alg == B200ENS_VERN7 ? 1024 : 4096)))) {
        auto keep_src = m->source;
        auto keep_cubin = m->cubin;
        auto keep_log = m->log;
        const int keep_spill = m->spill, keep_regs = m->regs, keep_lmem = m->lmem, keep_smem = m->smem;
        if (!try_split(keep_spill, true, split_on)) {
            m->source = keep_src;
            m->cubin = keep_cubin;
            m->log = keep_log;
            m->spill = keep_spill;
            m->regs = keep_regs;
            m->lmem = keep_lmem;
            m->smem = keep_smem;
        }
    }
    if (!rc && nvec && !m->split && (!try_regs || (m->spill > 4096 && !(force_k && atoi(force_k) == 0)))) {
        const int per_thread = nvec * d->n_state * (d->dtype == B200ENS_F64 ? 8 : 4);
        int block = std::min(128, (114688 / per_thread) / 32 * 32);
        if (block >= 32) {
            auto keep_src = m->source;
            auto keep_cubin = m->cubin;
            auto keep_log = m->log;
            const int keep_spill = m->spill, keep_regs = m->regs, keep_lmem = m->lmem, keep_smem = m->smem;
            m->source = build_source(d, 1, block, 1);
            int rc2 = nvrtc_compile(m.get());
            if (!rc2) parse_ptxas_log(m.get());
            if (!rc2 && (!try_regs || m->spill < keep_spill)) {
                m->ksmem = 1;
                m->block = block;
                m->kvec_bytes = per_thread;
                mb = 1;
            } else if (try_regs) {
                m->source = keep_src;
                m->cubin = keep_cubin;
                m->log = keep_log;
                m->spill = keep_spill;
                m->regs = keep_regs;
                m->lmem = keep_lmem;
                m->smem = keep_smem;
            } else {
                rc = rc2;
            }
        }
    }
    m->min_blocks = mb;
    if (log && log_len) {
        strncpy(log, m->log.c_str(), log_len - 1);
        log[log_len - 1] = 0;
    }
    if (rc) return rc;
    *out = m.release();
    return 0;
}

void b200ens_free(b200ens_model* m) {
    if (!m) return;
    if (m->lib) cudaLibraryUnload(m->lib);
    delete m;
}

int b200ens_model_info(const b200ens_model* m, int64_t* cubin_bytes, int32_t* regs, int32_t* smem, int32_t* lmem) {
    if (!m) return fail(B200ENS_E_INVALID, "null model");
    if (cubin_bytes) *cubin_bytes = (int64_t)m->cubin.size();
    if (regs) *regs = m->regs;
    if (smem) *smem = m->smem;
    if (lmem) *lmem = m->lmem;
    return 0;
}

static int solve_host(b200ens_model* m, const b200ens_opts* o, int64_t N, const void* u0, const void* p,
                      const void* saveat, int32_t n_save, const void* dW, void* out_u, void* out_t, int32_t* retcode,
                      b200ens_stats* stats, b200ens_timing* timing, double* msum, double* msumsq, int64_t* mcount) {
    const bool moments = msum != nullptr;
    int rc = check_opts(m, o, n_save, dW);
    if (rc) return rc;
    if (N < 0) return fail(B200ENS_E_INVALID, "N < 0");
    if (timing) memset(timing, 0, sizeof *timing);
    const bool every = o->save_everystep != 0;
    if (every && (moments || !out_t)) return fail(B200ENS_E_INVALID, "save_everystep needs out_t [N][n_save] (and no moments mode)");
    if (!u0 || !retcode || (m->n_param && !p) || (n_save && ((!saveat && !every) || (!out_u && !moments))))
        return fail(B200ENS_E_INVALID, "null buffer");
    if (n_save && !every) {
        // the header's contract (ascending, inside tspan): a point outside would silently stay unwritten.  Bounds are
        // compared in the state type, which is what the kernel sees (tspan (0, 0.1) in Float32 ends at 0.1f > 0.1).
        const bool f64 = m->dtype == B200ENS_F64;
        auto at = [&](int i) { return f64 ? ((const double*)saveat)[i] : (double)((const float*)saveat)[i]; };
        const double lo = f64 ? o->t0 : (double)(float)o->t0, hi = f64 ? o->t1 : (double)(float)o->t1;
        for (int i = 0; i < n_save; i++)
            if (!(at(i) >= lo && at(i) <= hi) || (i && !(at(i) >= at(i - 1))))
                return fail(B200ENS_E_INVALID, "saveat[%d] = %g: the grid must be ascending and inside tspan [%g, %g]", i, at(i), lo, hi);
    }
    if (out_t && n_save && !every) memcpy(out_t, saveat, (size_t)n_save * m->elem());
    if (N == 0) return 0;
    const int ndev = b200ens_device_count();
    if (ndev <= 0) return fail(B200ENS_E_NODEVICE, "no CUDA device available (libb200ens has no CPU fallback)");
    std::vector<int> devs;
    for (int g = 0; g < ndev && g < 32; g++)
        if (o->device_mask == 0 || (o->device_mask >> g) & 1u) devs.push_back(g);
    if (devs.empty()) return fail(B200ENS_E_NODEVICE, "device_mask 0x%x selects no visible device", o->device_mask);
    // tests: B200ENS_VIRTUAL_SHARDS=k deals the ensemble as if k devices were selected, all mapped onto the first one
    // (the shards then run one after the other) -- exercises the multi-device dealing and gather on a one-GPU box
    if (const char* e = getenv("B200ENS_VIRTUAL_SHARDS")) {
        const int k = std::max(1, std::min(32, atoi(e)));
        devs.assign((size_t)k, devs[0]);
    }
    if ((long long)devs.size() > N) devs.resize((size_t)N);
    const int G = (int)devs.size();
    std::vector<ShardResult> res(G);
    const size_t row_len = (size_t)n_save * m->n_out;
    std::vector<Moments> moms(G);
    std::vector<std::vector<double>> mbuf(G);
    if (moments)
        for (int g = 0; g < G; g++) {
            mbuf[g].assign(2 * row_len, 0.0);
            moms[g].sum = mbuf[g].data();
            moms[g].sumsq = mbuf[g].data() + row_len;
        }
    // Dealing the trajectories: the ensemble is cut into G*k contiguous blocks which go to the devices in
    // boustrophedon order 0,1,..,G-1,G-1,..,1,0,0,1,..: along an ORDERED parameter sweep the work per trajectory varies
    // ~10x (Lorenz rho-sweep), contiguous G-ths would leave the first GPU idle while the last one integrates the chaotic
    // end; the back-and-forth deal cancels a linear trend exactly and leaves every device the same mix.  Every block is
    // D2H-copied straight into its place in the caller's arrays -- the gather is still implicit, there is no collective
    // -- and every trajectory is computed exactly as before (Philox streams are keyed by the global index).
    int k_blocks = o->shard_blocks > 0 ? o->shard_blocks : 8;
    if (G == 1 || is_sde(m->alg)) k_blocks = 1;   // SDE steppers: uniform work, and Philox streams are keyed by consecutive indices
    while (k_blocks > 1 && N / ((long long)G * k_blocks) < 8192) k_blocks /= 2;   // pieces stay large enough for efficient copies
    std::vector<Ranges> deal(G);
    {
        const long long nb = (long long)G * k_blocks;
        for (long long b = 0; b < nb; b++) {
            const long long lo = N * b / nb, hi = N * (b + 1) / nb;
            const long long r = b % (2 * G);
            const int g = (int)(r < G ? r : 2 * G - 1 - r);
            if (hi > lo) deal[g].emplace_back(lo, hi);
        }
    }
    auto run = [&](int g) {
        res[g].code = solve_shard(m, o, devs[g], deal[g], (const char*)u0, (const char*)p, saveat, n_save,
                                  (const char*)dW, (char*)out_u, retcode, stats, &res[g], moments ? &moms[g] : nullptr,
                                  every ? (char*)out_t : nullptr);
        if (res[g].code) res[g].err = g_err;
    };
    if (G == 1) {
        run(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++) th.emplace_back(run, g);
        for (auto& t : th) t.join();
    }
    for (int g = 0; g < G; g++)
        if (res[g].code) {
            g_err = res[g].err;
            return res[g].code;
        }
    if (moments) {  // host gather of the per-device partial sums (no collective needed inside one process)
        for (size_t i = 0; i < row_len; i++) msum[i] = msumsq[i] = 0.0;
        *mcount = 0;
        for (int g = 0; g < G; g++) {
            for (size_t i = 0; i < row_len; i++) {
                msum[i] += moms[g].sum[i];
                msumsq[i] += moms[g].sumsq[i];
            }
            *mcount += moms[g].count;
        }
    }
    if (timing) {
        for (int g = 0; g < G; g++) {
            timing->h2d_ms = std::max(timing->h2d_ms, res[g].h2d);
            timing->kernel_ms = std::max(timing->kernel_ms, res[g].kern);
            timing->kernel_ms_min = g == 0 ? res[g].kern : std::min(timing->kernel_ms_min, res[g].kern);
            timing->d2h_ms = std::max(timing->d2h_ms, res[g].d2h);
            timing->total_ms = std::max(timing->total_ms, res[g].total);
            timing->launches += res[g].launches;
        }
        timing->n_devices = G;
        timing->grid = res[0].lp.grid;
        timing->block = res[0].lp.block;
        timing->smem_bytes = res[0].lp.smem;
        timing->regs = m->regs;
    }
    return 0;
}

int b200ens_solve(b200ens_model* m, const b200ens_opts* o, int64_t N, const void* u0, const void* p,
                  const void* saveat, int32_t n_save, const void* dW, void* out_u, void* out_t, int32_t* retcode,
                  b200ens_stats* stats, b200ens_timing* timing) {
    return solve_host(m, o, N, u0, p, saveat, n_save, dW, out_u, out_t, retcode, stats, timing, nullptr, nullptr, nullptr);
}

int b200ens_solve_moments(b200ens_model* m, const b200ens_opts* o, int64_t N, const void* u0, const void* p,
                          const void* saveat, int32_t n_save, const void* dW, double* sum, double* sumsq,
                          int64_t* count, int32_t* retcode, b200ens_timing* timing) {
    if (!sum || !sumsq || !count) return fail(B200ENS_E_INVALID, "null moments buffer");
    if (N == 0) {
        *count = 0;
        for (int i = 0; m && i < n_save * m->n_out; i++) sum[i] = sumsq[i] = 0.0;
    }
    return solve_host(m, o, N, u0, p, saveat, n_save, dW, nullptr, nullptr, retcode, nullptr, timing, sum, sumsq, count);
}

int b200ens_solve_device(b200ens_model* m, const b200ens_opts* o, int32_t device, void* stream, int64_t N,
                         const void* d_u0, const void* d_p, const void* d_saveat, int32_t n_save, const void* d_dW,
                         void* d_out_u, int32_t* d_retcode, b200ens_stats* d_stats, b200ens_timing* timing) {
    int rc = check_opts(m, o, n_save, d_dW);
    if (rc) return rc;
    if (timing) memset(timing, 0, sizeof *timing);
    if (o->save_everystep) return fail(B200ENS_E_UNSUPPORTED, "save_everystep is served by b200ens_solve (host buffers) only");
    if (N <= 0) return N == 0 ? 0 : fail(B200ENS_E_INVALID, "N < 0");
    if (!d_u0 || !d_retcode || (m->n_param && !d_p) || (n_save && (!d_saveat || !d_out_u)))
        return fail(B200ENS_E_INVALID, "null buffer");
    const int ndev = b200ens_device_count();
    if (ndev <= 0) return fail(B200ENS_E_NODEVICE, "no CUDA device available (libb200ens has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(B200ENS_E_INVALID, "device %d not visible", device);
    DeviceCtx* d;
    if ((rc = device_ctx(device, &d))) return rc;
    CU(cudaSetDevice(device));
    if ((rc = ensure_loaded(m))) return rc;
    B2Args a{};
    if ((rc = fill_args(m, o, &a))) return rc;
    LaunchPlan lp;
    if ((rc = plan_launch(m, o, d, N, n_save, &lp))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* counter = d->counters + (d->ring.fetch_add(1) % kRing);
    a.u0 = d_u0;
    a.p = d_p;
    a.saveat = d_saveat;
    a.dW = d_dW;
    a.out_u = d_out_u;
    a.retcode = d_retcode;
    a.stats = d_stats;
    a.work_counter = counter;
    a.N = N;
    a.traj_offset = o->traj_offset;
    a.n_save = n_save;
    a.refill_threshold = lp.refill;
    a.stage_stride = lp.stride;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (timing) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        CU(cudaEventRecord(e0, st));
    }
    void* d_ts = nullptr;
    {
        int n_ts = 0;
        const std::vector<char> ts_host = tstops_bytes(m, o, &n_ts);
        if (n_ts) {   // pageable source: the copy is staged before cudaMemcpyAsync returns
            CU(cudaMallocAsync(&d_ts, ts_host.size(), st));
            CU(cudaMemcpyAsync(d_ts, ts_host.data(), ts_host.size(), cudaMemcpyHostToDevice, st));
            CU(cudaStreamSynchronize(st));
        }
        a.tstops = d_ts;
        a.n_tstops = n_ts;
    }
    void* work = nullptr;
    int launches = 1;
    if (want_work_order(m, o, a, N)) {
        // stream-ordered scratch: after warm-up the pool hands the same block back without touching the driver
        CU(cudaMallocAsync(&work, work_scratch_bytes(N), st));
        if ((rc = enqueue_work_order(m, o, d, &a, work, st))) return rc;
        launches += 2;
    } else {
        CU(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
    }
    if ((rc = launch(m, lp, a, st))) return rc;
    if (work) CU(cudaFreeAsync(work, st));
    if (d_ts) CU(cudaFreeAsync(d_ts, st));
    if (timing) {
        CU(cudaEventRecord(e1, st));
        CU(cudaEventSynchronize(e1));
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        timing->kernel_ms = ms;
        timing->total_ms = ms;
        timing->n_devices = 1;
        timing->launches = launches;
        timing->grid = lp.grid;
        timing->block = lp.block;
        timing->smem_bytes = lp.smem;
        timing->regs = m->regs;
        CU(cudaGetLastError());
    }
    return 0;
}

const char* b200ens_nvrtc_info(void) {
    static thread_local std::string info;
    const NvrtcApi* nv = nvrtc_api();
    if (!nv) return "";
    info = std::to_string(nv->major) + "." + std::to_string(nv->minor) + " " + nv->path;
    return info.c_str();
}

void* b200ens_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        (void)cudaGetLastError();
        fail(B200ENS_E_NOMEM, "cudaHostAlloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void b200ens_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

}  // extern "C"
