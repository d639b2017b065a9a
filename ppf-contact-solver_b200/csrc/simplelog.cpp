// simplelog.cpp -- the per-phase log side channel of the backend.
//
// The frontend reads `<data_dir>/<scope>.<name>.out` files ("time value" lines) by name
// (crates/ppf-cts-core/src/parsers.rs:6-53), so the file names and line formats are part of the
// drop-in contract.  This restates the behaviour of the reference's SimpleLog
// (crates/ppf-cts-solver/src/cpp/simplelog/SimpleLog.cpp:86-178): scope total written on exit,
// push/pop timers in milliseconds, mark() values (integers as %d, others as %e).
#include "backend.hpp"
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <map>
#include <stack>
#include <sys/stat.h>

namespace ppf {

static std::string g_dir;
static double g_time = 0.0;
static int g_depth = 0;
static bool g_quiet = getenv("PPF_B200_QUIET") != nullptr;

static std::string norm_name(const char *n) {
    std::string s(n);
    for (auto &c : s)
        if (c == ' ') c = '_';
    return s;
}

struct ScopeLog::Impl {
    std::string name;
    std::chrono::steady_clock::time_point start;
    std::stack<std::pair<std::chrono::steady_clock::time_point, std::string>> stack;
    std::map<std::string, std::vector<double>> dict;
};

void simplelog_set_path(const char *dir) {
    g_dir = dir ? dir : "";
    if (!g_dir.empty()) mkdir(g_dir.c_str(), 0755);
}
void simplelog_set_time(double t) { g_time = t; }

ScopeLog::ScopeLog(const char *name) : impl(new Impl) {
    impl->name = norm_name(name);
    impl->start = std::chrono::steady_clock::now();
    ++g_depth;
    message("====== %s ======", impl->name.c_str());
}

ScopeLog::~ScopeLog() {
    const auto end = std::chrono::steady_clock::now();
    const unsigned long ms = (unsigned long)std::chrono::duration_cast<std::chrono::milliseconds>(end - impl->start).count();
    message("===== %s: %lu msec =====", impl->name.c_str(), ms);
    if (!g_dir.empty()) {
        FILE *fp = fopen((g_dir + "/" + impl->name + ".out").c_str(), "a");
        if (fp) { fprintf(fp, "%f %lu\n", g_time, ms); fclose(fp); }
        for (auto &it : impl->dict) {
            fp = fopen((g_dir + "/" + impl->name + "." + it.first + ".out").c_str(), "a");
            if (!fp) continue;
            for (double e : it.second) {
                if (!fmodf((float)e, 1.0f)) fprintf(fp, "%f %d\n", g_time, (int)e);
                else fprintf(fp, "%f %e\n", g_time, e);
            }
            fclose(fp);
        }
    }
    --g_depth;
    delete impl;
}

void ScopeLog::push(const char *name) { impl->stack.push({std::chrono::steady_clock::now(), norm_name(name)}); }

void ScopeLog::pop() {
    if (impl->stack.empty()) return;
    const std::string name = impl->stack.top().second;
    const auto end = std::chrono::steady_clock::now();
    const unsigned long ms = (unsigned long)std::chrono::duration_cast<std::chrono::milliseconds>(end - impl->stack.top().first).count();
    impl->stack.pop();
    message("> %s...%lu msec", name.c_str(), ms);
    if (!g_dir.empty()) {
        FILE *fp = fopen((g_dir + "/" + impl->name + "." + name + ".out").c_str(), "a");
        if (fp) { fprintf(fp, "%f %lu\n", g_time, ms); fclose(fp); }
    }
}

void ScopeLog::mark(const char *name, double value) {
    impl->dict[norm_name(name)].push_back(value);
    if (!fmodf((float)value, 1.0f)) message("* %s: %d", name, (int)value);
    else message("* %s: %.3e", name, value);
}

void ScopeLog::message(const char *fmt, ...) {
    if (g_quiet) return;
    for (int i = 1; i < g_depth; ++i) printf("   ");
    va_list ap;
    va_start(ap, fmt);
    vprintf(fmt, ap);
    va_end(ap);
    printf("\n");
    fflush(stdout);
}

} // namespace ppf
