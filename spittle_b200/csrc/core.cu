// Library-wide state: thread-local last error, version, launch counter.
#include "common.cuh"
#include <atomic>
#include <cstring>
#include <cstdlib>

namespace sb {
static thread_local std::string t_last_error;
std::atomic<uint64_t> g_launches{0};
static bool pdl_from_env() { const char* e = getenv("SB_PDL"); return !(e && e[0] == '0'); }
bool g_pdl = pdl_from_env();
thread_local TraceSlot g_trace_next;

void set_error(const std::string& msg) { t_last_error = msg; }
const char* get_error() { return t_last_error.c_str(); }
}  // namespace sb

extern "C" {
const char* sb_last_error(void) { return sb::get_error(); }
const char* sb_version(void) { return "spittle_b200 0.1.0 (sm_100a)"; }
uint64_t sb_launch_count(void) { return sb::g_launches.load(); }
}
