// Host-side input staging (include/adni_staging.h): gzip-compressed NIfTI-1 volumes -> fp32 / uint8 buffers in the
// memory order of torch.tensor(nib.load(path).get_fdata()), decoded by a pool of threads straight into the caller's
// (pinned) batch buffer.  Restates nibabel 4.0.2 (the reference's reader, environment.yml:154; absent from the
// reference tree) for single-file NIfTI-1 images; call sites replaced: pkg/utils/dataloader.py:206-207, 226-227,
// 239-241.  No CUDA in this library.
//
// Built with -ffp-contract=off: get_fdata computes raw * slope and + inter as two rounded fp64 operations.
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/adni_staging.h"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

template <typename T>
T bswap(T v) {
  unsigned char b[sizeof(T)];
  std::memcpy(b, &v, sizeof(T));
  for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
  std::memcpy(&v, b, sizeof(T));
  return v;
}
template <typename T>
T rd(const unsigned char* p, bool swap) {
  T v;
  std::memcpy(&v, p, sizeof(T));
  return swap ? bswap(v) : v;
}

int bytes_of(int datatype) {
  switch (datatype) {
    case 2: case 256: return 1;
    case 4: case 512: return 2;
    case 8: case 768: case 16: return 4;
    case 64: case 1024: case 1280: return 8;
    default: return 0;
  }
}

struct GzFile {
  gzFile f = nullptr;
  explicit GzFile(const char* path) { f = gzopen(path, "rb"); if (f) gzbuffer(f, 1 << 18); }
  ~GzFile() { if (f) gzclose(f); }
  bool read(void* dst, size_t n) {
    unsigned char* p = static_cast<unsigned char*>(dst);
    while (n) {
      const unsigned chunk = n > (1u << 30) ? (1u << 30) : static_cast<unsigned>(n);
      const int got = gzread(f, p, chunk);
      if (got <= 0) return false;
      p += got;
      n -= static_cast<size_t>(got);
    }
    return true;
  }
  bool skip(size_t n) {
    unsigned char tmp[4096];
    while (n) {
      const size_t c = n > sizeof tmp ? sizeof tmp : n;
      if (!read(tmp, c)) return false;
      n -= c;
    }
    return true;
  }
};

// NIfTI-1 header (348 bytes): sizeof_hdr@0 i32, dim@40 i16[8], datatype@70 i16, bitpix@72 i16, vox_offset@108 f32,
// scl_slope@112 f32, scl_inter@116 f32, magic@344 char[4] ("n+1\0" single file, "ni1\0" header/image pair).
int parse_header(const unsigned char* h, const char* path, adni_nifti_info* info) {
  std::memset(info, 0, sizeof *info);
  int32_t sz = rd<int32_t>(h, false);
  bool swap = false;
  if (sz != 348) {
    if (bswap(sz) == 348) swap = true;
    else if (sz == 540 || bswap(sz) == 540) return fail(ADNI_STAGE_ENOTSUP, "%s: NIfTI-2 images are not supported", path);
    else return fail(ADNI_STAGE_ENOTSUP, "%s: not a NIfTI-1 file (sizeof_hdr = %d)", path, sz);
  }
  if (std::memcmp(h + 344, "n+1", 3) != 0) {
    if (std::memcmp(h + 344, "ni1", 3) == 0) return fail(ADNI_STAGE_ENOTSUP, "%s: header/image pairs (.hdr/.img) are not supported", path);
    return fail(ADNI_STAGE_ENOTSUP, "%s: bad NIfTI-1 magic", path);
  }
  int16_t dim[8];
  for (int i = 0; i < 8; i++) dim[i] = rd<int16_t>(h + 40 + 2 * i, swap);
  if (dim[0] < 1 || dim[0] > 7) return fail(ADNI_STAGE_ENOTSUP, "%s: dim[0] = %d out of range", path, dim[0]);
  int nd = dim[0];
  int64_t nvox = 1;
  for (int i = 0; i < 7; i++) info->dim[i] = 1;
  for (int i = 0; i < nd; i++) {
    if (dim[i + 1] < 0) return fail(ADNI_STAGE_ENOTSUP, "%s: negative extent", path);
    info->dim[i] = dim[i + 1];
    // checked product: seven int16 extents can reach 2^105, and a wrapped value could pass the bound below
    if (__builtin_mul_overflow(nvox, static_cast<int64_t>(dim[i + 1]), &nvox) || nvox > (int64_t(1) << 40))
      return fail(ADNI_STAGE_ENOTSUP, "%s: implausible voxel count", path);
  }
  info->ndim = nd;
  info->nvox = nvox;
  info->datatype = rd<int16_t>(h + 70, swap);
  info->bitpix = rd<int16_t>(h + 72, swap);
  if (!bytes_of(info->datatype)) return fail(ADNI_STAGE_ENOTSUP, "%s: datatype code %d is not supported", path, info->datatype);
  const float vox = rd<float>(h + 108, swap);
  info->vox_offset = (vox >= 352.f && vox < 1e9f) ? static_cast<int64_t>(vox) : 352;  // single-file images start at >= 352
  const double slope = static_cast<double>(rd<float>(h + 112, swap));
  const double inter = static_cast<double>(rd<float>(h + 116, swap));
  info->swapped = swap ? 1 : 0;
  // nibabel Nifti1Header.get_slope_inter: slope 0 or non-finite -> no scaling; a non-finite intercept is an error
  if (slope == 0.0 || !std::isfinite(slope)) {
    info->scaled = 0;
    info->scl_slope = 1.0;
    info->scl_inter = 0.0;
  } else {
    if (!std::isfinite(inter)) return fail(ADNI_STAGE_ENOTSUP, "%s: scl_inter is not finite", path);
    info->scaled = (slope != 1.0 || inter != 0.0) ? 1 : 0;
    info->scl_slope = slope;
    info->scl_inter = inter;
  }
  return ADNI_STAGE_OK;
}

int open_and_parse(const char* path, GzFile& gz, adni_nifti_info* info) {
  if (!path || !info) return fail(ADNI_STAGE_EINVAL, "null argument");
  if (!gz.f) return fail(ADNI_STAGE_EIO, "%s: cannot open", path);
  unsigned char h[348];
  if (!gz.read(h, sizeof h)) return fail(ADNI_STAGE_EIO, "%s: truncated header", path);
  return parse_header(h, path, info);
}

template <typename T>
inline double scaled_value(const unsigned char* p, bool swap, bool scaled, double slope, double inter) {
  double v = static_cast<double>(rd<T>(p, swap));
  if (scaled) {
    v = v * slope;  // two rounded operations, as numpy evaluates arr * slope + inter
    v = v + inter;
  }
  return v;
}

// OUT = double | float | uint8 mask.  Disk order: axis 0 fastest.  Memory order: axis 0 slowest (C order of the
// (d0, d1, d2[, ...]) array).  The transposition walks tiles of kT axis-0 elements so that reads are contiguous runs
// and writes are kT sequential streams.
template <typename T, typename OUT>
int convert(const unsigned char* raw, const adni_nifti_info& in, OUT* dst, int* nonbinary) {
  const bool swap = in.swapped, scaled = in.scaled;
  const double slope = in.scl_slope, inter = in.scl_inter;
  const int64_t d0 = in.dim[0];
  int64_t rest = in.nvox / (d0 ? d0 : 1);  // product of the remaining axes, in disk order (axis 1 fastest)
  // C-order index of the "rest" part: reverse the axis order.  For 3-D volumes: disk r = j + d1*k  ->  mem j*d2 + k.
  const int nd = in.ndim;
  std::vector<int64_t> memidx(static_cast<size_t>(rest));
  {
    int64_t ext[6], stride[6];
    for (int a = 1; a < nd; a++) ext[a - 1] = in.dim[a];
    int64_t s = 1;
    for (int a = nd - 1; a >= 1; a--) {
      stride[a - 1] = s;
      s *= in.dim[a];
    }
    int64_t idx[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t r = 0; r < rest; r++) {
      int64_t m = 0;
      for (int a = 0; a < nd - 1; a++) m += idx[a] * stride[a];
      memidx[static_cast<size_t>(r)] = m;
      for (int a = 0; a < nd - 1; a++) {
        if (++idx[a] < ext[a]) break;
        idx[a] = 0;
      }
    }
  }
  constexpr int kT = 16;
  int bad = 0;
  for (int64_t i0 = 0; i0 < d0; i0 += kT) {
    const int64_t ni = (d0 - i0 < kT) ? (d0 - i0) : kT;
    for (int64_t r = 0; r < rest; r++) {
      const unsigned char* p = raw + (r * d0 + i0) * static_cast<int64_t>(sizeof(T));
      const int64_t m = memidx[static_cast<size_t>(r)];
      for (int64_t i = 0; i < ni; i++) {
        const double v = scaled_value<T>(p + i * sizeof(T), swap, scaled, slope, inter);
        OUT* o = dst + (i0 + i) * rest + m;
        if (nonbinary) {
          if (v != 0.0 && v != 1.0) bad++;
          *o = static_cast<OUT>(v != 0.0 ? 1 : 0);
        } else {
          *o = static_cast<OUT>(v);
        }
      }
    }
  }
  if (nonbinary) *nonbinary = bad;
  return ADNI_STAGE_OK;
}

template <typename OUT>
int convert_any(const unsigned char* raw, const adni_nifti_info& in, OUT* dst, int* nonbinary) {
  switch (in.datatype) {
    case 2: return convert<uint8_t, OUT>(raw, in, dst, nonbinary);
    case 256: return convert<int8_t, OUT>(raw, in, dst, nonbinary);
    case 4: return convert<int16_t, OUT>(raw, in, dst, nonbinary);
    case 512: return convert<uint16_t, OUT>(raw, in, dst, nonbinary);
    case 8: return convert<int32_t, OUT>(raw, in, dst, nonbinary);
    case 768: return convert<uint32_t, OUT>(raw, in, dst, nonbinary);
    case 16: return convert<float, OUT>(raw, in, dst, nonbinary);
    case 64: return convert<double, OUT>(raw, in, dst, nonbinary);
    case 1024: return convert<int64_t, OUT>(raw, in, dst, nonbinary);
    case 1280: return convert<uint64_t, OUT>(raw, in, dst, nonbinary);
  }
  return ADNI_STAGE_ENOTSUP;
}

template <typename OUT>
int read_volume_impl(const char* path, OUT* dst, int64_t capacity, adni_nifti_info* info_out, bool mask);

// No exception may cross the C ABI (or escape a worker thread): allocation failures become an error code.
template <typename OUT>
int read_volume(const char* path, OUT* dst, int64_t capacity, adni_nifti_info* info_out, bool mask) {
  try {
    return read_volume_impl<OUT>(path, dst, capacity, info_out, mask);
  } catch (const std::exception& e) {
    return fail(ADNI_STAGE_EIO, "%s: %s", path ? path : "(null)", e.what());
  } catch (...) {
    return fail(ADNI_STAGE_EIO, "%s: unknown failure", path ? path : "(null)");
  }
}

template <typename OUT>
int read_volume_impl(const char* path, OUT* dst, int64_t capacity, adni_nifti_info* info_out, bool mask) {
  adni_nifti_info info;
  GzFile gz(path ? path : "");
  int rc = open_and_parse(path, gz, &info);
  if (rc) return rc;
  if (info_out) *info_out = info;
  if (!dst) return fail(ADNI_STAGE_EINVAL, "%s: null destination", path);
  if (capacity < info.nvox) return fail(ADNI_STAGE_EINVAL, "%s: %lld voxels do not fit a buffer of %lld", path,
                                        static_cast<long long>(info.nvox), static_cast<long long>(capacity));
  if (!gz.skip(static_cast<size_t>(info.vox_offset - 348))) return fail(ADNI_STAGE_EIO, "%s: truncated before the image data", path);
  const size_t nbytes = static_cast<size_t>(info.nvox) * static_cast<size_t>(bytes_of(info.datatype));
  std::vector<unsigned char> raw(nbytes);
  if (nbytes && !gz.read(raw.data(), nbytes)) return fail(ADNI_STAGE_EIO, "%s: truncated or corrupt image data", path);
  int nonbinary = 0;
  rc = convert_any<OUT>(raw.data(), info, dst, mask ? &nonbinary : nullptr);
  if (rc) return fail(rc, "%s: datatype code %d is not supported", path, info.datatype);
  if (mask && nonbinary)
    return fail(ADNI_STAGE_ENOTSUP, "%s: %d mask voxels are neither 0 nor 1 (the reference multiplies by the mask)", path, nonbinary);
  return ADNI_STAGE_OK;
}

}  // namespace

extern "C" {

const char* adni_stage_last_error(void) { return g_err.c_str(); }
int adni_stage_version(void) { return 1; }

int adni_nifti_read_info(const char* path, adni_nifti_info* info) {
  GzFile gz(path ? path : "");
  return open_and_parse(path, gz, info);
}

int adni_nifti_read_f64(const char* path, double* dst, int64_t capacity, adni_nifti_info* info) {
  return read_volume<double>(path, dst, capacity, info, false);
}

int adni_nifti_read_f32(const char* path, float* dst, int64_t capacity, adni_nifti_info* info) {
  return read_volume<float>(path, dst, capacity, info, false);
}

int adni_nifti_read_mask_u8(const char* path, uint8_t* dst, int64_t capacity, adni_nifti_info* info) {
  return read_volume<uint8_t>(path, dst, capacity, info, true);
}

int adni_stage_volumes(const char* const* paths, int n, int kind, void* dst, int64_t stride_bytes, int64_t expect_nvox,
                       int threads, int* status) {
  if (!paths || n < 0 || !dst || (kind != 0 && kind != 1) || expect_nvox <= 0)
    return fail(ADNI_STAGE_EINVAL, "stage_volumes: bad arguments");
  const int64_t elem = kind == 0 ? 4 : 1;
  if (stride_bytes < expect_nvox * elem) return fail(ADNI_STAGE_EINVAL, "stage_volumes: stride smaller than one volume");
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  // Nothing may cross the C ABI: allocation of the status vectors, thread creation (std::system_error) and the
  // error-string copies are all inside the try block; threads that did start are always joined.
  std::vector<std::thread> pool;
  int first = 0;
  try {
    std::vector<int> codes(static_cast<size_t>(n), 0);
    std::vector<std::string> errs(static_cast<size_t>(n));
    std::atomic<int> next(0);
    auto work = [&]() {
      for (;;) {
        const int i = next.fetch_add(1);
        if (i >= n) break;
        if (!paths[i]) continue;
        int rc;
        try {
          adni_nifti_info info;
          char* slot = static_cast<char*>(dst) + static_cast<int64_t>(i) * stride_bytes;
          rc = kind == 0 ? read_volume<float>(paths[i], reinterpret_cast<float*>(slot), expect_nvox, &info, false)
                         : read_volume<uint8_t>(paths[i], reinterpret_cast<uint8_t*>(slot), expect_nvox, &info, true);
          if (rc == 0 && info.nvox != expect_nvox)
            rc = fail(ADNI_STAGE_EINVAL, "%s: %lld voxels, the batch expects %lld", paths[i],
                      static_cast<long long>(info.nvox), static_cast<long long>(expect_nvox));
          if (rc) errs[static_cast<size_t>(i)] = g_err;  // g_err is thread-local: carry it to the caller's thread
        } catch (...) {
          rc = ADNI_STAGE_EIO;  // (the string copy above may itself have thrown: keep the code, drop the text)
        }
        codes[static_cast<size_t>(i)] = rc;
      }
    };
    try {
      for (int t = 1; t < threads; t++) pool.emplace_back(work);
    } catch (...) {
      // could not start every helper: the ones running plus this thread still drain the queue
    }
    work();
    for (auto& th : pool) th.join();
    pool.clear();
    for (int i = 0; i < n; i++) {
      if (status) status[i] = codes[static_cast<size_t>(i)];
      if (!first && codes[static_cast<size_t>(i)]) {
        first = codes[static_cast<size_t>(i)];
        if (!errs[static_cast<size_t>(i)].empty()) g_err = errs[static_cast<size_t>(i)];
      }
    }
  } catch (...) {
    for (auto& th : pool)
      if (th.joinable()) th.join();
    return fail(ADNI_STAGE_EIO, "stage_volumes: out of memory or thread failure");
  }
  return first;
}

}  // extern "C"
