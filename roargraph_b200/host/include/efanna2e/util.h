// File loaders and small helpers with the reference's names and behaviour (include/efanna2e/util.h of the reference):
//   load_meta :107-127, load_data :180-211 (rows padded to a multiple of 8 floats, zero filled), data_align :37-75,
//   load_gt_meta :85-105 (the truthset must carry distances: ids block + distance block), load_gt_data_with_dist
//   :130-155, load_gt_data :157-176, normalize :214-225.
// Differences: whole-file block reads instead of one ifstream::read per row, and a dimension that is not a multiple
// of 8 is handled consistently (the reference's data_align re-reads already padded rows with the unpadded stride,
// which is only correct when dim % 8 == 0 — SURVEY.md 9.10).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>
#if defined(__SSE__) || defined(__x86_64__)
#include <xmmintrin.h>
#endif

namespace efanna2e {

constexpr unsigned kDataAlignFactor = 8;

inline size_t file_size_or_die(std::ifstream &in) {
  if (!in.is_open()) {
    std::cout << "open file error" << std::endl;
    exit(-1);
  }
  in.seekg(0, std::ios::end);
  return (size_t)in.tellg();
}

template <typename T>
void load_meta(const char *filename, unsigned &points_num, unsigned &dim) {
  std::ifstream in(filename, std::ios::binary);
  const size_t fsize = file_size_or_die(in);
  in.seekg(0, std::ios::beg);
  in.read((char *)&points_num, 4);
  in.read((char *)&dim, 4);
  const uint32_t contained = dim ? (uint32_t)((fsize - 8) / dim / sizeof(T)) : 0;
  std::cout << "load meta from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
  if (points_num != contained) {
    std::cerr << "filename: " << filename << std::endl;
    std::cerr << "Data file size wrong! Get points " << contained << " but should have " << points_num << std::endl;
    throw std::runtime_error("Data file size wrong!");
  }
}

template <typename T>
void load_gt_meta(const char *filename, unsigned &points_num, unsigned &dim) {
  std::ifstream in(filename, std::ios::binary);
  const size_t fsize = file_size_or_die(in);
  in.seekg(0, std::ios::beg);
  in.read((char *)&points_num, 4);
  in.read((char *)&dim, 4);
  const uint32_t contained = dim ? (uint32_t)((fsize - 8) / dim / sizeof(T)) : 0;
  std::cout << "load gt from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
  if (points_num * 2 != contained) {  // ids block followed by an equally sized distance block
    std::cerr << "filename: " << filename << std::endl;
    std::cerr << "Data file size wrong! Get points " << contained << " but should have " << points_num << std::endl;
    throw std::runtime_error("Data file size wrong!");
  }
}

// Rows land with stride = dim rounded up to a multiple of 8, tail zero filled; `dim` is NOT updated (as in the
// reference) — data_align reports the padded value.
template <typename T>
void load_data(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data) {
  std::cout << "load data from file: " << filename << std::endl;
  std::ifstream in(filename, std::ios::binary);
  const size_t fsize = file_size_or_die(in);
  const uint64_t pts = points_num, d = dim;
  if (fsize != pts * d * sizeof(T) + 8) {
    std::cerr << "Read file incompleted! filename:" << filename << std::endl;
    throw std::runtime_error("Data file size wrong!");
  }
  const uint64_t nd = (d + kDataAlignFactor - 1) / kDataAlignFactor * kDataAlignFactor;
  if (posix_memalign((void **)&data, 64, pts * nd * sizeof(T)) != 0) throw std::bad_alloc();
  in.seekg(8, std::ios::beg);
  if (nd == d) {
    in.read((char *)data, (std::streamsize)(pts * d * sizeof(T)));
  } else {
    std::vector<T> row(d);
    for (uint64_t i = 0; i < pts; ++i) {
      in.read((char *)row.data(), (std::streamsize)(d * sizeof(T)));
      memcpy(data + i * nd, row.data(), d * sizeof(T));
      memset(data + i * nd + d, 0, (nd - d) * sizeof(T));
    }
  }
  std::cout << "Finish load data from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
}

// Takes rows laid out by load_data (stride = padded dim) and returns them as the aligned block the index uses;
// updates `dim` to the padded value.  Ownership of data_ori passes to the returned pointer (free() it).
inline float *data_align(float *data_ori, unsigned point_num, unsigned &dim) {
  (void)point_num;
  dim = (dim + kDataAlignFactor - 1) / kDataAlignFactor * kDataAlignFactor;
  std::cout << "new_dim: " << dim << std::endl;
  return data_ori;
}

template <typename T, typename T2>
void load_gt_data_with_dist(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data, T2 *&res_dists) {
  std::ifstream in(filename, std::ios::binary);
  const size_t fsize = file_size_or_die(in);
  const size_t n = (size_t)points_num * dim;
  if (fsize != n * sizeof(T) + n * sizeof(T2) + 8) {
    std::cerr << "Read file incompleted!" << std::endl;
    throw std::runtime_error("Data file size wrong!");
  }
  in.seekg(8, std::ios::beg);
  data = new T[n];
  res_dists = new T2[n];
  in.read((char *)data, (std::streamsize)(n * sizeof(T)));
  in.read((char *)res_dists, (std::streamsize)(n * sizeof(T2)));
}

template <typename T>
void load_gt_data(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data) {
  std::ifstream in(filename, std::ios::binary);
  const size_t fsize = file_size_or_die(in);
  const size_t n = (size_t)points_num * dim;
  if (fsize < n * sizeof(T) + 8) {
    std::cerr << "Read file incompleted!" << std::endl;
    throw std::runtime_error("Data file size wrong!");
  }
  in.seekg(8, std::ios::beg);
  data = new T[n];
  in.read((char *)data, (std::streamsize)(n * sizeof(T)));
}

// normalize<T> (:214-225 of the reference): x / |x| in place.  The reference's release build (-Ofast -march=native)
// compiles the loop to 8 lane accumulators with separately rounded multiply and add, the folds (l+4)+l, (l+2)+l, 1+0,
// RSQRTSS plus one Newton-Raphson step inv = (r*(r*sum) - 3) * (-0.5*r) with the inner product fused, and a MULTIPLY by
// inv.  For dim % 8 == 0 on x86 this reproduces that sequence (RSQRTSS is executed, not modelled), so cosine inputs reach
// the device bit-identical to what the reference's loader / CLI hand to its own search; otherwise the source-level
// arithmetic is used (equal to rounding).
namespace detail {
#if defined(__SSE__) || defined(__x86_64__)
inline bool normalize_as_compiled(float *arr, const size_t dim) {
  if (dim == 0 || dim % 8 != 0) return false;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (size_t i = 0; i < dim; i += 8)
    for (size_t l = 0; l < 8; ++l) {
      volatile float prod = arr[i + l] * arr[i + l];
      acc[l] = acc[l] + prod;
    }
  const float x0 = acc[4] + acc[0], x1 = acc[5] + acc[1], x2 = acc[6] + acc[2], x3 = acc[7] + acc[3];
  const float y0 = x2 + x0, y1 = x3 + x1;
  const float sum = y1 + y0;
  const float rs = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(sum)));
  volatile float t = rs * sum;
  t = std::fmaf(rs, t, -3.0f);
  const float r2 = rs * -0.5f;
  const float inv = t * r2;
  for (size_t i = 0; i < dim; ++i) arr[i] = arr[i] * inv;
  return true;
}
#else
inline bool normalize_as_compiled(float *, const size_t) { return false; }
#endif
template <typename T>
inline bool normalize_fast(T *, const size_t) { return false; }
template <>
inline bool normalize_fast<float>(float *arr, const size_t dim) { return normalize_as_compiled(arr, dim); }
}  // namespace detail

template <typename T>
inline void normalize(T *arr, const size_t dim) {
  if (detail::normalize_fast<T>(arr, dim)) return;
  float sum = 0.0f;
  for (size_t i = 0; i < dim; i++) sum += arr[i] * arr[i];
  sum = std::sqrt(sum);
  for (size_t i = 0; i < dim; i++) arr[i] = (T)(arr[i] / sum);
}

// normalize() over n rows of stride `dim` on `threads` host threads (0 = all cores); every row is independent, so the
// result does not depend on the thread count.  The reference runs this loop serially at load time
// (src/index_bipartite.cpp:2679-2684): 10 M x 512 rows take it ~10 s.
inline void normalize_rows(float *rows, size_t n, size_t dim, unsigned threads = 0) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  threads = (unsigned)std::min<size_t>(threads, std::max<size_t>(1, n / 4096));
  if (threads <= 1) {
    for (size_t i = 0; i < n; ++i) normalize<float>(rows + i * dim, dim);
    return;
  }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < threads; ++t)
    pool.emplace_back([=] {
      for (size_t i = n * t / threads, e = n * (t + 1) / threads; i < e; ++i) normalize<float>(rows + i * dim, dim);
    });
  for (auto &th : pool) th.join();
}

}  // namespace efanna2e
