// shader.cu — shader objects of the B200 device: IR -> CUDA source (ir_lower.cpp) -> NVRTC (sm_100a cubin) -> cudaLibrary.
//
// The counterpart of the reference CPU backend's cpu/shader.rs (clang++ on the generated C++, dlopen of the result) and of
// ShaderImpl::new (cpu/shader.rs:150-258): compile once per create_shader, keep the entry point, launch per ShaderDispatch with
// dispatch_size clipped per thread (cpu/stream.rs:330-440).  NVRTC is loaded lazily with dlopen so that the library itself
// has no link-time dependency on it; a missing libnvrtc fails loudly at the first create_shader.
#include "shader.h"
#include "trace_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace lcb {

namespace {

// the three headers a generated kernel includes, embedded at build time (Makefile: embedded_headers.inc)
struct EmbeddedHeader { const char *name; const char *text; };
#include "embedded_headers.inc"

typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    int (*CompileProgram)(nvrtcProgram, int, const char *const *);
    int (*GetProgramLogSize)(nvrtcProgram, size_t *);
    int (*GetProgramLog)(nvrtcProgram, char *);
    int (*GetCUBINSize)(nvrtcProgram, size_t *);
    int (*GetCUBIN)(nvrtcProgram, char *);
    int (*DestroyProgram)(nvrtcProgram *);
    const char *(*GetErrorString)(int);
};

const Nvrtc &nvrtc() {
    static Nvrtc api = [] {
        // The traversal header uses 256-bit global loads (PTX ISA 8.8): NVRTC older than 12.9 cannot assemble them.  A process
        // that imported PyTorch already holds torch's bundled (older) libnvrtc.so.12 under that soname, so the toolkit's copy is
        // looked up by path first and every candidate is checked for its version.
        const char *candidates[] = {getenv("LC_B200_NVRTC"), "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/targets/x86_64-linux/lib/libnvrtc.so.12",
                                    "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so.12", "libnvrtc.so"};
        void *h = nullptr;
        std::string seen;
        for (const char *c : candidates) {
            if (!c || h) continue;
            void *lib = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (!lib) continue;
            int major = 0, minor = 0;
            auto version = (int (*)(int *, int *))dlsym(lib, "nvrtcVersion");
            if (version && version(&major, &minor) == 0 && (major > 12 || (major == 12 && minor >= 9))) h = lib;
            else { seen += std::string(" ") + c + " (" + std::to_string(major) + "." + std::to_string(minor) + ")"; dlclose(lib); }
        }
        if (!h) throw std::runtime_error("create_shader needs NVRTC >= 12.9 (libnvrtc.so.12) and none could be loaded; tried:" + seen + "; set LC_B200_NVRTC to its path");
        Nvrtc a{};
        auto sym = [&](const char *n) { void *p = dlsym(h, n); if (!p) throw std::runtime_error(std::string("NVRTC symbol missing: ") + n); return p; };
        a.CreateProgram = (decltype(a.CreateProgram))sym("nvrtcCreateProgram");
        a.CompileProgram = (decltype(a.CompileProgram))sym("nvrtcCompileProgram");
        a.GetProgramLogSize = (decltype(a.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
        a.GetProgramLog = (decltype(a.GetProgramLog))sym("nvrtcGetProgramLog");
        a.GetCUBINSize = (decltype(a.GetCUBINSize))sym("nvrtcGetCUBINSize");
        a.GetCUBIN = (decltype(a.GetCUBIN))sym("nvrtcGetCUBIN");
        a.DestroyProgram = (decltype(a.DestroyProgram))sym("nvrtcDestroyProgram");
        a.GetErrorString = (decltype(a.GetErrorString))sym("nvrtcGetErrorString");
        return a;
    }();
    return api;
}

}  // namespace

struct ShaderObj {
    LoweredKernel lowered;
    std::vector<char> cubin;
    cudaLibrary_t library = nullptr;
    cudaKernel_t kernel = nullptr;
    unsigned resident_ctas = 0;  // wavefront-lowered kernels: SMs x CTAs per SM (occupancy query at the first launch)
    std::string name;
};

ShaderObj *shader_create(const ir::KernelModule *km, bool fast_math, bool compile_only, const char *name, std::string &log) {
    auto *s = new ShaderObj;
    try {
        lower_kernel(km, s->lowered);
        if (name) s->name = name;
        if (const char *dump = getenv("LC_B200_DUMP_KERNELS")) {
            static int serial = 0;
            char path[512]; snprintf(path, sizeof(path), "%s/lc_kernel_%d.cu", dump, serial++);
            if (FILE *f = fopen(path, "w")) { fputs(s->lowered.source.c_str(), f); fclose(f); }
        }
        // identical source + options -> identical cubin: kernels re-created in a process (the frontend's enable_cache) skip NVRTC
        static std::mutex cache_mu;
        static std::unordered_map<std::string, std::vector<char>> cache;
        const std::string cache_key = (fast_math ? "F" : "P") + s->lowered.source;
        bool cached = false;
        {
            std::lock_guard<std::mutex> lk(cache_mu);
            auto it = cache.find(cache_key);
            if (it != cache.end()) { s->cubin = it->second; cached = true; }
        }
        if (!cached) {
        const Nvrtc &rt = nvrtc();
        const size_t n_headers = sizeof(kEmbeddedHeaders) / sizeof(kEmbeddedHeaders[0]);
        std::vector<const char *> names, texts;
        for (size_t i = 0; i < n_headers; i++) { names.push_back(kEmbeddedHeaders[i].name); texts.push_back(kEmbeddedHeaders[i].text); }
        nvrtcProgram prog = nullptr;
        int rc = rt.CreateProgram(&prog, s->lowered.source.c_str(), "lc_kernel.cu", (int)n_headers, texts.data(), names.data());
        if (rc != 0) throw std::runtime_error(std::string("nvrtcCreateProgram: ") + rt.GetErrorString(rc));
        // fp32 semantics: no FMA contraction, IEEE division and square root, denormals kept — the reference CPU backend compiles
        // without fast-math (cpu/shader.rs:41-45) and results must not depend on the optimiser's contraction choices.
        std::vector<const char *> opts = {"-arch=sm_100a", "-std=c++17", "-default-device", "-lineinfo", "-diag-suppress=177"};
        if (fast_math) opts.push_back("-use_fast_math");
        else { opts.push_back("-fmad=false"); opts.push_back("-prec-div=true"); opts.push_back("-prec-sqrt=true"); opts.push_back("-ftz=false"); }
        rc = rt.CompileProgram(prog, (int)opts.size(), opts.data());
        size_t log_size = 0;
        rt.GetProgramLogSize(prog, &log_size);
        if (log_size > 1) { log.resize(log_size); rt.GetProgramLog(prog, &log[0]); }
        if (rc != 0) {
            rt.DestroyProgram(&prog);
            throw std::runtime_error(std::string("NVRTC failed to compile the lowered kernel (") + rt.GetErrorString(rc) + "):\n" + log);
        }
        size_t sz = 0;
        rt.GetCUBINSize(prog, &sz);
        s->cubin.resize(sz);
        rt.GetCUBIN(prog, s->cubin.data());
        rt.DestroyProgram(&prog);
        std::lock_guard<std::mutex> lk(cache_mu);
        if (cache.size() < 256) cache[cache_key] = s->cubin;
        }
        if (!compile_only) {
            cudaError_t e = cudaLibraryLoadData(&s->library, s->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
            if (e != cudaSuccess) throw std::runtime_error(std::string("cudaLibraryLoadData: ") + cudaGetErrorString(e));
            e = cudaLibraryGetKernel(&s->kernel, s->library, "lc_kernel");
            if (e != cudaSuccess) throw std::runtime_error(std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e));
        }
    } catch (...) {
        delete s;
        throw;
    }
    return s;
}

void shader_destroy(ShaderObj *s) {
    if (!s) return;
    if (s->library) cudaLibraryUnload(s->library);
    delete s;
}

const LoweredKernel &shader_lowered(const ShaderObj *s) { return s->lowered; }

// LC_B200_WAVE_YIELD overrides the per-kernel choice of ir_lower.cpp (read per launch: a tuning knob for sweeps)
static int wave_yield_min(int chosen) {
    const char *e = getenv("LC_B200_WAVE_YIELD");
    const int y = e ? atoi(e) : chosen;
    return y < 1 ? 1 : (y > 32 ? 32 : y);
}

int shader_launch(ShaderObj *s, cudaStream_t stream, void *params, const uint32_t dispatch_size[3], unsigned long long *work_counter) {
    if (!s->kernel) throw std::runtime_error("shader was created with compile_only and cannot be dispatched");
    const uint32_t *b = s->lowered.block_size;
    if (dispatch_size[0] == 0 || dispatch_size[1] == 0 || dispatch_size[2] == 0) return 0;
    dim3 grid((dispatch_size[0] + b[0] - 1) / b[0], (dispatch_size[1] + b[1] - 1) / b[1], (dispatch_size[2] + b[2] - 1) / b[2]);
    dim3 block(b[0], b[1], b[2]);
    if (s->lowered.wave) {
        // persistent grid: every resident thread slot pulls dispatch ids (in the block-major order of the grid above) from the counter
        if (s->resident_ctas == 0) {
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)s->kernel, kWaveThreads, 0) != cudaSuccess || per_sm < 1) { (void)cudaGetLastError(); per_sm = 1; }
            s->resident_ctas = (unsigned)(sms * per_sm);
        }
        HostLaunch *launch = reinterpret_cast<HostLaunch *>(params);
        launch->work_items = (unsigned long long)grid.x * grid.y * grid.z * b[0] * b[1] * b[2];
        launch->work_counter = work_counter;
        launch->yield_min = (uint32_t)wave_yield_min(s->lowered.wave_yield_min);
        const unsigned long long want = (launch->work_items + kWaveThreads - 1) / kWaveThreads;
        grid = dim3((unsigned)(want < s->resident_ctas ? want : s->resident_ctas), 1, 1);
        block = dim3(kWaveThreads, 1, 1);
        cudaError_t e = cudaMemsetAsync(work_counter, 0, sizeof(unsigned long long), stream);
        if (e != cudaSuccess) throw std::runtime_error(std::string("work counter reset failed: ") + cudaGetErrorString(e));
    }
    void *args[1] = {params};
    cudaError_t e = cudaLaunchKernel((const void *)s->kernel, grid, block, args, 0, stream);
    if (e != cudaSuccess) throw std::runtime_error(std::string("kernel launch failed: ") + cudaGetErrorString(e));
    return 1;
}

}  // namespace lcb
