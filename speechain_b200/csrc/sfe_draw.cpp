// SpecAugment draws on torch's default CPU generator in ONE call (host glue, not on the device path).
//
// The reference draws with torch.randint on the CPU generator, in a fixed order, with bounds that depend on the
// previous draw's maximum (specaug.py:98-108,145-158,178-191).  The module layer reproduces that sequence so that a
// seeded run masks exactly what the reference's CPU path masks; done from Python it is four to six tensor-op
// dispatches and two reductions per step (~10 us).  This extension makes the same draws from the same generator —
// at::CPUGeneratorImpl::random(), mapped to [low, high) exactly like ATen's random_from_to kernel
// (uniform_int_from_to: (random() % range) + low, random64() for ranges of 2^32 and more), serially in memory order —
// and writes them as int32 straight into the persistent descriptor arrays the kernels read.
#include <ATen/CPUGeneratorImpl.h>
#include <torch/extension.h>

#include <mutex>
#include <tuple>

namespace {

inline int64_t draw_one(at::CPUGeneratorImpl* gen, int64_t low, int64_t high) {
  TORCH_CHECK(low < high, "random_ expects 'from' to be less than 'to', but got from=", low, " >= to=", high);
  const uint64_t range = static_cast<uint64_t>(high - low);
  if (range >= (1ULL << 32)) return static_cast<int64_t>(gen->random64() % range) + low;
  return static_cast<int64_t>(gen->random() % range) + low;
}

// fills out[0..n) and returns the maximum
inline int64_t draw_many(at::CPUGeneratorImpl* gen, int64_t low, int64_t high, int32_t* out, int64_t n) {
  int64_t mx = INT64_MIN;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t v = draw_one(gen, low, high);
    out[i] = static_cast<int32_t>(v);
    if (v > mx) mx = v;
  }
  return mx;
}

}  // namespace

// Returns (warp_center, warp_pos), (-1, -1) without a warp.  Pointers are host addresses of (B, nf) / (B, nt) int32
// arrays (0 where the kind of mask is off).  t_lo / t_hi: the time-mask width bounds already resolved against minlen.
std::tuple<int64_t, int64_t> specaug_draw(int64_t B, int64_t feat_dim, int64_t minlen, int64_t warp_window,
                                          int64_t nf, int64_t f_lo, int64_t f_hi, int64_t fpos_ptr, int64_t flen_ptr,
                                          int64_t nt, int64_t t_lo, int64_t t_hi, int64_t tpos_ptr, int64_t tlen_ptr) {
  auto* gen = at::get_generator_or_default<at::CPUGeneratorImpl>(c10::nullopt, at::detail::getDefaultCPUGenerator());
  std::lock_guard<std::mutex> lock(gen->mutex_);
  int64_t center = -1, pos = -1;
  if (warp_window >= 0 && minlen > 2 * warp_window + 1) {
    center = draw_one(gen, warp_window + 1, minlen - warp_window);
    pos = draw_one(gen, center - warp_window, center + warp_window);
  }
  if (nf > 0) {
    const int64_t mx = draw_many(gen, f_lo, f_hi + 1, reinterpret_cast<int32_t*>(flen_ptr), B * nf);
    draw_many(gen, 0, std::max<int64_t>(1, feat_dim - mx), reinterpret_cast<int32_t*>(fpos_ptr), B * nf);
  }
  if (nt > 0) {
    const int64_t mx = draw_many(gen, t_lo, t_hi + 1, reinterpret_cast<int32_t*>(tlen_ptr), B * nt);
    draw_many(gen, 0, std::max<int64_t>(1, minlen - mx), reinterpret_cast<int32_t*>(tpos_ptr), B * nt);
  }
  return std::make_tuple(center, pos);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("specaug_draw", &specaug_draw, "SpecAugment draws on the default CPU generator (see sfe_draw.cpp)");
}
