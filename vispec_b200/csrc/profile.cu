#include "profile.h"

#include <atomic>
#include <vector>

#include "../../include/vispec_b200.h"

namespace vs {
namespace {
std::atomic<long long> g_launches{0};
struct Rec { cudaEvent_t a, b; int cls; double bytes; };
bool g_on = false;
size_t g_cap = 0;
std::vector<Rec> g_recs;
bool g_open = false;
}  // namespace

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

void prof_begin(int cls, double bytes, cudaStream_t st) {
  if (!g_on || g_recs.size() >= g_cap) return;
  Rec r;
  r.cls = cls;
  r.bytes = bytes;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_recs.push_back(r);
  g_open = true;
}
void prof_end(cudaStream_t st) {
  if (!g_open) return;
  cudaEventRecord(g_recs.back().b, st);
  g_open = false;
}
}  // namespace vs

extern "C" {
long long vs_launch_count(int reset) {
  long long v = vs::g_launches.load();
  if (reset) vs::g_launches.store(0);
  return v;
}
int vs_profile_start(int max_records) {
  vs::g_recs.clear();
  vs::g_recs.reserve(max_records > 0 ? max_records : 0);
  vs::g_cap = max_records > 0 ? (size_t)max_records : 0;
  vs::g_on = true;
  return VS_OK;
}
int vs_profile_stop(float* ms_out, double* bytes_out, int32_t* cls_out, int cap) {
  vs::g_on = false;
  cudaDeviceSynchronize();
  int n = 0;
  for (auto& r : vs::g_recs) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (n < cap) {
      ms_out[n] = ms;
      bytes_out[n] = r.bytes;
      cls_out[n] = r.cls;
      ++n;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  vs::g_recs.clear();
  return n;
}
}
