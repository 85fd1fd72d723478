// Host-side plumbing of libavzoom: error string, constant tables, argument checks.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <vector>

#include "avz_common.cuh"

namespace avz {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct TableEntry {
  int device;
  int n_fft;
  Tables t;
};
static std::mutex g_mu;
static std::vector<TableEntry> g_tables;

int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

int tables_for(int n_fft, Tables* out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(AVZ_ENOGPU, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_mu);
  for (const auto& te : g_tables)
    if (te.device == dev && te.n_fft == n_fft) {
      *out = te.t;
      return AVZ_OK;
    }
  const int n = n_fft;
  std::vector<double2> twd(n);
  std::vector<double> wd(n);
  std::vector<float2> twf(n);
  std::vector<float> wf(n);
  const double two_pi = 6.283185307179586476925286766559;
  for (int k = 0; k < n; ++k) {
    // exact values at the quadrant points so that +-1 / +-i twiddles carry no rounding noise
    double c = cos(two_pi * k / n), s = -sin(two_pi * k / n);
    if ((4 * k) % n == 0) {
      const int q = (4 * k) / n;
      c = (q == 0) ? 1.0 : (q == 2 ? -1.0 : 0.0);
      s = (q == 1) ? -1.0 : (q == 3 ? 1.0 : 0.0);
    }
    twd[k] = make_double2(c, s);
    twf[k] = make_float2((float)c, (float)s);
    wd[k] = 0.5 - 0.5 * cos(two_pi * k / n);  // scipy get_window('hann', n) (periodic)
    wf[k] = (float)wd[k];
  }
  void *p_tw = nullptr, *p_w = nullptr, *p_twd = nullptr, *p_wd = nullptr;
  AVZ_CUDA_OK(cudaMalloc(&p_tw, n * sizeof(float2)));
  AVZ_CUDA_OK(cudaMalloc(&p_w, n * sizeof(float)));
  AVZ_CUDA_OK(cudaMalloc(&p_twd, n * sizeof(double2)));
  AVZ_CUDA_OK(cudaMalloc(&p_wd, n * sizeof(double)));
  AVZ_CUDA_OK(cudaMemcpy(p_tw, twf.data(), n * sizeof(float2), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(p_w, wf.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(p_twd, twd.data(), n * sizeof(double2), cudaMemcpyHostToDevice));
  AVZ_CUDA_OK(cudaMemcpy(p_wd, wd.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  TableEntry te;
  te.device = dev;
  te.n_fft = n_fft;
  te.t.tw = (const float2*)p_tw;
  te.t.win = (const float*)p_w;
  te.t.tw_d = (const double2*)p_twd;
  te.t.win_d = (const double*)p_wd;
  g_tables.push_back(te);
  *out = te.t;
  return AVZ_OK;
}

static bool g_prof_on = false;
static cudaEvent_t g_prof_ev[PROF_COUNT][2];
static bool g_prof_made = false;
static bool g_prof_hit[PROF_COUNT];

void prof_begin(int slot, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof_ev[slot][0], st);
}
void prof_end(int slot, cudaStream_t st) {
  if (!g_prof_on) return;
  cudaEventRecord(g_prof_ev[slot][1], st);
  g_prof_hit[slot] = true;
}

int check_fft_args(int n_fft, int hop, int64_t L) {
  if (n_fft != 256 && n_fft != 512 && n_fft != 1024)
    return set_error(AVZ_EINVAL, "n_fft=%d unsupported (256, 512, 1024)", n_fft);
  if (hop <= 0 || n_fft % hop != 0 || n_fft / hop < 2 || n_fft / hop > 8)
    return set_error(AVZ_EINVAL, "hop=%d unsupported for n_fft=%d (need n_fft %% hop == 0, 2 <= n_fft/hop <= 8)", hop,
                     n_fft);
  if (L < n_fft)
    return set_error(AVZ_EINVAL, "L=%lld shorter than n_fft=%d (scipy would shrink nperseg; not supported)",
                     (long long)L, n_fft);
  return AVZ_OK;
}

}  // namespace avz

extern "C" {

int avz_version(void) { return AVZ_VERSION; }

const char* avz_last_error(void) { return avz::g_err; }

#define AVZ_STR2(x) #x
#define AVZ_STR(x) AVZ_STR2(x)
const char* avz_build_info(void) {
  return "libavzoom " AVZ_STR(AVZ_VERSION) " sm_100a nvcc " AVZ_STR(__CUDACC_VER_MAJOR__) "." AVZ_STR(__CUDACC_VER_MINOR__) "." AVZ_STR(
      __CUDACC_VER_BUILD__) " -O3 -lineinfo "
#ifdef AVZ_EXPERIMENT
         "experiment (environment knobs enabled)";
#else
         "release (no environment reads)";
#endif
}

int avz_init(int n_fft) {
  if (n_fft != 256 && n_fft != 512 && n_fft != 1024) return avz::set_error(AVZ_EINVAL, "n_fft=%d unsupported", n_fft);
  avz::Tables t;
  if (n_fft == 1024) {   // the 1024 / hop 512 fast path runs on the 512-point transform: it needs both tables
    const int rc = avz::tables_for(512, &t);
    if (rc) return rc;
  }
  return avz::tables_for(n_fft, &t);
}

int avz_profile_enable(int on) {
  using namespace avz;
  if (on && !g_prof_made) {
    for (int i = 0; i < PROF_COUNT; ++i)
      for (int j = 0; j < 2; ++j) AVZ_CUDA_OK(cudaEventCreate(&g_prof_ev[i][j]));
    g_prof_made = true;
  }
  for (int i = 0; i < PROF_COUNT; ++i) g_prof_hit[i] = false;
  g_prof_on = on != 0;
  return AVZ_OK;
}

int avz_profile_get(float* ms_host, int n) {
  using namespace avz;
  if (!ms_host || n < PROF_COUNT) return set_error(AVZ_EINVAL, "avz_profile_get: need room for %d floats", (int)PROF_COUNT);
  for (int i = 0; i < PROF_COUNT; ++i) {
    ms_host[i] = -1.f;
    if (g_prof_made && g_prof_hit[i]) {
      AVZ_CUDA_OK(cudaEventSynchronize(g_prof_ev[i][1]));
      AVZ_CUDA_OK(cudaEventElapsedTime(&ms_host[i], g_prof_ev[i][0], g_prof_ev[i][1]));
    }
  }
  return AVZ_OK;
}

int64_t avz_num_frames(int64_t L, int n_fft, int hop) {
  if (L <= 0 || n_fft <= 0 || hop <= 0) return 0;
  const int64_t ext = L + 2 * (int64_t)(n_fft / 2);
  int64_t r = (ext - n_fft) % hop;          // >= 0 since ext >= n_fft
  int64_t nadd = ((hop - r) % hop) % n_fft; // (-(ext - n) mod hop) mod n
  return (ext + nadd - n_fft) / hop + 1;
}

}  // extern "C"
