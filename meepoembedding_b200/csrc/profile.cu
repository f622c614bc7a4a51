// profile.cu — optional per-kernel CUDA-event timing on the launch stream (bench.py's roofline leg).
// Off by default: when off, ProfScope costs one branch and records nothing.
#include <cstdio>
#include <cstring>
#include <map>

#include "table.h"

namespace meepo {

struct ProfRecord {
  const char* name;
  cudaEvent_t a, b;
};
struct Profiler {
  bool on = false;
  std::vector<ProfRecord> pending;
  std::vector<cudaEvent_t> pool;
  std::map<std::string, std::pair<uint64_t, double>> acc;  // name -> (launches, total ms)
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void drain() {
    for (auto& r : pending) {
      float ms = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        auto& x = acc[r.name];
        x.first += 1;
        x.second += ms;
      }
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    pending.clear();
  }
};

ProfScope::ProfScope(meepo_table* t, const char* name, cudaStream_t s) : t_(t), name_(name), s_(s) {
  if (t_->prof && t_->prof->on) {
    a_ = t_->prof->get();
    cudaEventRecord(a_, s_);
  }
}
ProfScope::~ProfScope() {
  if (a_) {
    cudaEvent_t b = t_->prof->get();
    cudaEventRecord(b, s_);
    t_->prof->pending.push_back(ProfRecord{name_, a_, b});
  }
}
// host-side phases (wall clock) reported next to the kernels
void prof_add_host(meepo_table* t, const char* name, double ms) {
  if (!t->prof || !t->prof->on) return;
  auto& x = t->prof->acc[name];
  x.first += 1;
  x.second += ms;
}
void destroy_profiler(meepo_table* t) {
  if (!t->prof) return;
  t->prof->drain();
  for (auto e : t->prof->pool) cudaEventDestroy(e);
  delete t->prof;
  t->prof = nullptr;
}

}  // namespace meepo

using namespace meepo;

extern "C" {

MEEPO_API meepo_status meepo_profile_enable(meepo_table* t, int32_t on) {
  if (!t) return fail(MEEPO_EINVAL, "null table");
  DeviceGuard guard(t->device);
  if (!t->prof) t->prof = new Profiler();
  t->prof->drain();
  if (on) t->prof->acc.clear();
  t->prof->on = on != 0;
  return MEEPO_OK;
}

MEEPO_API meepo_status meepo_profile_read(meepo_table* t, char* buf, uint64_t buf_bytes) {
  if (!t || !buf || !buf_bytes) return fail(MEEPO_EINVAL, "null argument");
  DeviceGuard guard(t->device);
  std::string out;
  if (t->prof) {
    t->prof->drain();
    char line[256];
    for (auto& kv : t->prof->acc) {
      snprintf(line, sizeof line, "%s %llu %.6f\n", kv.first.c_str(), (unsigned long long)kv.second.first,
               kv.second.second);
      out += line;
    }
  }
  if (out.size() + 1 > buf_bytes) return fail(MEEPO_EINVAL, "profile buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return MEEPO_OK;
}

}  // extern "C"
