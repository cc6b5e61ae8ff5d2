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
#include <algorithm>
#include <cerrno>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
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

namespace {

// SHA-256 (FIPS 180-4), for the kernel cache key: the reference caches compiled kernels by the SHA-256 of their source
// (cpu/shader.rs:84-152: <exe dir>/.cache/kernel_<hash>.bc).
struct Sha256 {
    uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint8_t block[64]; size_t fill = 0; uint64_t total = 0;
    static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
    void compress(const uint8_t *p) {
        static const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu,
            0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u,
            0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, 0xa2bfe8a1u, 0xa81a664bu,
            0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
            0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[64];
        for (int i = 0; i < 16; i++) w[i] = (uint32_t)p[4 * i] << 24 | (uint32_t)p[4 * i + 1] << 16 | (uint32_t)p[4 * i + 2] << 8 | p[4 * i + 3];
        for (int i = 16; i < 64; i++) {
            const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16] + s0 + w[i - 7] + s1;
        }
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
        for (int i = 0; i < 64; i++) {
            const uint32_t t1 = hh + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g)) + K[i] + w[i];
            const uint32_t t2 = (rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    void update(const void *data, size_t n) {
        const uint8_t *p = (const uint8_t *)data;
        total += n;
        while (n) {
            const size_t k = std::min(n, 64 - fill);
            memcpy(block + fill, p, k); fill += k; p += k; n -= k;
            if (fill == 64) { compress(block); fill = 0; }
        }
    }
    std::string hex() {
        const uint64_t bits = total * 8;
        const uint8_t one = 0x80, zero = 0;
        update(&one, 1);
        while (fill != 56) update(&zero, 1);
        uint8_t len[8];
        for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(len, 8);
        char out[65];
        for (int i = 0; i < 8; i++) snprintf(out + 8 * i, 9, "%08x", h[i]);
        return std::string(out, 64);
    }
};

// <directory of this library>/.cache (the reference: next to the executable), LC_B200_CACHE_DIR overrides, LC_B200_CACHE=0 disables
std::string cache_dir() {
    static std::string dir = [] {
        if (const char *off = getenv("LC_B200_CACHE")) if (off[0] == '0') return std::string();
        std::string d;
        if (const char *e = getenv("LC_B200_CACHE_DIR")) d = e;
        else {
            Dl_info info{};
            if (!dladdr((const void *)&cache_dir, &info) || !info.dli_fname) return std::string();
            d = info.dli_fname;
            const size_t slash = d.find_last_of('/');
            d = (slash == std::string::npos ? std::string(".") : d.substr(0, slash)) + "/.cache";
        }
        if (mkdir(d.c_str(), 0755) != 0 && errno != EEXIST) return std::string();
        return d;
    }();
    return dir;
}

bool read_file(const std::string &path, std::vector<char> &out) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); const long n = ftell(f); fseek(f, 0, SEEK_SET);
    if (n <= 0) { fclose(f); return false; }
    out.resize((size_t)n);
    const bool ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

void write_file_atomically(const std::string &path, const std::vector<char> &data) {
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
    fclose(f);
    if (!ok || rename(tmp.c_str(), path.c_str()) != 0) remove(tmp.c_str());
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
        // identical source + options -> identical cubin: kernels re-created in a process (the frontend's enable_cache) skip NVRTC ...
        static std::mutex cache_mu;
        static std::unordered_map<std::string, std::vector<char>> cache;
        const std::string cache_key = (fast_math ? "F" : "P") + s->lowered.source;
        bool cached = false;
        {
            std::lock_guard<std::mutex> lk(cache_mu);
            auto it = cache.find(cache_key);
            if (it != cache.end()) { s->cubin = it->second; cached = true; }
        }
        // ... and across processes: <cache dir>/kernel_<SHA-256 of options + device library headers + source>.cubin
        std::string disk_path;
        if (!cached && !cache_dir().empty()) {
            Sha256 sha;
            sha.update(fast_math ? "F" : "P", 1);
            const size_t n_hdr = sizeof(kEmbeddedHeaders) / sizeof(kEmbeddedHeaders[0]);
            for (size_t i = 0; i < n_hdr; i++) sha.update(kEmbeddedHeaders[i].text, strlen(kEmbeddedHeaders[i].text));
            sha.update(s->lowered.source.data(), s->lowered.source.size());
            disk_path = cache_dir() + "/kernel_" + sha.hex() + ".cubin";
            if (read_file(disk_path, s->cubin)) {
                cached = true;
                std::lock_guard<std::mutex> lk(cache_mu);
                if (cache.size() < 256) cache[cache_key] = s->cubin;
            }
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
        // enable_fast_math (the frontend's default, runtime/kernel.rs:564; the reference's CUDA backend maps it to --use_fast_math): approximate
        // division / square root / transcendentals and FMA contraction in the user's arithmetic.  Denormals are kept (-ftz=false): the
        // traversal's canonical triangle arithmetic is written with explicit round-to-nearest intrinsics, which contraction and the
        // approximate operators leave alone but which a global flush-to-zero would change — hits stay the same bits under fast math.
        if (fast_math) { opts.push_back("-use_fast_math"); opts.push_back("-ftz=false"); }
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
        if (!disk_path.empty()) write_file_atomically(disk_path, s->cubin);
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
