// rx_jitc.cpp -- rxjitc, the compiler process of the batch-shader JIT (rx_jit.cu).
//
// NVRTC runs here and not inside the host application: a compilation takes seconds, the library compiles in the background, and
// a process that exits while NVRTC is busy on another thread tears NVRTC's global state down under it (exit() runs the
// destructors of statics NVRTC creates lazily; no atexit ordering of ours can get in front of those).  In a child process the
// host can simply kill the compilation; it also keeps NVRTC's LLVM symbols and ~100 MB of code out of the host's address space.
//
// usage: rxjitc <workdir> <name expression> <output file> [option ...]
//   <workdir>/rx_kernels.cu is the translation unit, every other file of <workdir> an include (-I <workdir>);
//   output: u32 length of the lowered kernel name, the name, the cubin (written to <output>.tmp, then renamed);
//   the compiler's messages go to <output>.log; exit status 0 = compiled.
#include <dlfcn.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

typedef struct _nvrtcProgram* nvrtcProgram;

static bool read_file(const std::string& path, std::string* out) {
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) out->append(buf, got);
    fclose(fp);
    return true;
}

static void write_log(const std::string& out, const std::string& text) {
    if (FILE* fp = fopen((out + ".log").c_str(), "wb")) { fwrite(text.data(), 1, text.size(), fp); fclose(fp); }
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: rxjitc <workdir> <name expression> <output file> [nvrtc option ...]\n"); return 2; }
    const std::string dir = argv[1], name_expr = argv[2], out = argv[3];
    void* so = nullptr;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    for (const char* nm : names) if (!so) so = dlopen(nm, RTLD_NOW);
    if (!so) { write_log(out, "libnvrtc.so.12 not found"); return 3; }
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*DestroyProgram)(nvrtcProgram*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*AddNameExpression)(nvrtcProgram, const char*);
    int (*GetLoweredName)(nvrtcProgram, const char*, const char**);
    bool ok = true;
#define RX_SYM(f, n) *(void**)(&f) = dlsym(so, n); if (!f) ok = false
    RX_SYM(CreateProgram, "nvrtcCreateProgram"); RX_SYM(DestroyProgram, "nvrtcDestroyProgram"); RX_SYM(CompileProgram, "nvrtcCompileProgram");
    RX_SYM(GetCUBINSize, "nvrtcGetCUBINSize"); RX_SYM(GetCUBIN, "nvrtcGetCUBIN"); RX_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
    RX_SYM(GetProgramLog, "nvrtcGetProgramLog"); RX_SYM(AddNameExpression, "nvrtcAddNameExpression"); RX_SYM(GetLoweredName, "nvrtcGetLoweredName");
#undef RX_SYM
    if (!ok) { write_log(out, "libnvrtc lacks a required entry point"); return 3; }

    std::string src;
    if (!read_file(dir + "/rx_kernels.cu", &src)) { write_log(out, "no rx_kernels.cu in " + dir); return 2; }
    nvrtcProgram prog = nullptr;
    if (CreateProgram(&prog, src.c_str(), "rx_kernels.cu", 0, nullptr, nullptr) != 0) { write_log(out, "nvrtcCreateProgram failed"); return 1; }
    AddNameExpression(prog, name_expr.c_str());
    const std::string inc = "-I" + dir;
    std::vector<const char*> opts = {inc.c_str()};
    for (int i = 4; i < argc; ++i) opts.push_back(argv[i]);
    const int rc = CompileProgram(prog, (int)opts.size(), opts.data());
    std::string log;
    size_t ls = 0;
    if (GetProgramLogSize(prog, &ls) == 0 && ls > 1) { log.resize(ls); GetProgramLog(prog, &log[0]); }
    const char* low = nullptr;
    size_t cs = 0;
    if (rc != 0 || GetLoweredName(prog, name_expr.c_str(), &low) != 0 || !low || GetCUBINSize(prog, &cs) != 0 || cs == 0) {
        write_log(out, log.empty() ? std::string("nvrtcCompileProgram failed") : log);
        DestroyProgram(&prog);
        return 1;
    }
    std::vector<char> cubin(cs);
    GetCUBIN(prog, cubin.data());
    const std::string lowered = low;
    DestroyProgram(&prog);
    const std::string tmp = out + ".tmp";
    FILE* fp = fopen(tmp.c_str(), "wb");
    if (!fp) { write_log(out, "cannot write " + tmp); return 1; }
    const uint32_t nlen = (uint32_t)lowered.size();
    const bool wrote = fwrite(&nlen, 4, 1, fp) == 1 && fwrite(lowered.data(), 1, nlen, fp) == nlen && fwrite(cubin.data(), 1, cs, fp) == cs;
    fclose(fp);
    if (!wrote || rename(tmp.c_str(), out.c_str()) != 0) { remove(tmp.c_str()); write_log(out, "cannot write " + out); return 1; }
    if (!log.empty()) write_log(out, log);
    return 0;
}
