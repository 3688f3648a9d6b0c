// extern "C" boundary of libagpt_b200.so (declared in include/agpt_b200.h).
#include <atomic>
#include "models.h"

namespace agpt {
static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};
void set_last_error(const std::string& msg) { g_last_error = msg; }
void count_launch(long n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

template <typename F>
static int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 1;
  } catch (...) {
    set_last_error("unknown C++ exception");
    return 2;
  }
}

static Handle* as(agpt_handle h, uint32_t magic, const char* what) {
  auto* p = reinterpret_cast<Handle*>(h);
  if (!p || p->magic != magic) throw Error(std::string("invalid handle: expected ") + what);
  return p;
}
}  // namespace agpt

using namespace agpt;

extern "C" {

const char* agpt_last_error(void) { return g_last_error.c_str(); }
int agpt_version(void) { return 100; }
long long agpt_launch_count(void) { return g_launches.load(); }

void agpt_destroy(agpt_handle h) {
  auto* p = reinterpret_cast<Handle*>(h);
  if (!p) return;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  p->magic = 0;
  delete p;
}

int agpt_hifigan_create(const agpt_hifigan_cfg* cfg, const float* const* host_weights, int n_weights,
                        int device, agpt_handle* out) {
  return guarded([&] {
    AGPT_CHECK(cfg && host_weights && out, "null argument");
    *out = reinterpret_cast<agpt_handle>(hifigan_create(cfg, host_weights, n_weights, device));
  });
}

int agpt_hifigan_forward(agpt_handle h, const float* mel, const float* har_source, int B, int T, float* wav,
                         void* stream) {
  return guarded([&] {
    AGPT_CHECK(mel && wav, "null tensor");
    hifigan_forward(as(h, kMagicHifigan, "hifigan"), mel, har_source, B, T, wav, (cudaStream_t)stream);
  });
}

int agpt_hifigan_vocode_host(agpt_handle h, const float* mel_host, const float* har_host, int B, int T,
                             float* wav_host) {
  return guarded([&] {
    AGPT_CHECK(mel_host && wav_host, "null tensor");
    hifigan_vocode_host(as(h, kMagicHifigan, "hifigan"), mel_host, har_host, B, T, wav_host);
  });
}

}  // extern "C"
