// Trajectory ingestion, native side: DCD frames -> [count][N][3] float nm in a
// caller-owned (pinned) buffer.
//
// A DCD frame stores X, Y and Z as three separate float32 records in Angstrom (the layout OpenMM's
// DCDReporter writes); the kernels want interleaved nm.  The Python reader did this with strided
// numpy stores in a thread pool (7.7k frames/s at 100k atoms, a quarter of what PCIe can take).
// Here the file is mapped once, frames are split over host threads, and every thread converts
// whole frames with a loop the compiler vectorises: 3 contiguous reads, one interleaved write.
// Host code only -- no kernel is launched from this file.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

struct cvp_dcd {
  int fd = -1;
  const unsigned char* map = nullptr;
  size_t size = 0;
  int64_t num_atoms = 0, num_frames = 0, data_offset = 0, frame_bytes = 0;
  int has_box = 0;
  int32_t first_step = 0, interval = 0;
};

namespace cvp {
namespace {

int32_t read_i32(const unsigned char* p) {
  int32_t v;
  std::memcpy(&v, p, 4);
  return v;
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" {

int cvp_dcd_open(const char* path, cvp_dcd** out) {
  CVP_REQUIRE(path != nullptr && out != nullptr, "null argument");
  auto* f = new cvp_dcd();
  auto fail = [&](const std::string& why) {
    set_error(std::string(path) + ": " + why);
    if (f->map) munmap(const_cast<unsigned char*>(f->map), f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
    return CVP_ERR_INVALID;
  };
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) return fail("cannot open file");
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 100) return fail("not a DCD file (too short)");
  f->size = (size_t)st.st_size;
  void* map = mmap(nullptr, f->size, PROT_READ, MAP_SHARED, f->fd, 0);
  if (map == MAP_FAILED) return fail("mmap failed");
  f->map = static_cast<const unsigned char*>(map);
  const unsigned char* p = f->map;
  if (read_i32(p) != 84 || std::memcmp(p + 4, "CORD", 4) != 0 || read_i32(p + 88) != 84)
    return fail("not a little-endian CORD DCD file");
  int32_t icntrl[20];
  std::memcpy(icntrl, p + 8, 80);
  f->first_step = icntrl[1];
  f->interval = icntrl[2];
  f->has_box = icntrl[10] == 1;
  if (icntrl[11] == 1) return fail("4-dimensional DCD files are not supported");
  if (icntrl[8] != 0) return fail("fixed-atom (X-PLOR) DCD files are not supported");
  size_t pos = 92;
  if (pos + 4 > f->size) return fail("corrupt DCD title block");
  const int32_t title_bytes = read_i32(p + pos);
  pos += 4;
  if (title_bytes < 0 || pos + (size_t)title_bytes + 4 > f->size || read_i32(p + pos + title_bytes) != title_bytes)
    return fail("corrupt DCD title block");
  pos += (size_t)title_bytes + 4;
  if (pos + 12 > f->size || read_i32(p + pos) != 4 || read_i32(p + pos + 8) != 4 || read_i32(p + pos + 4) <= 0)
    return fail("corrupt DCD atom-count block");
  f->num_atoms = read_i32(p + pos + 4);
  pos += 12;
  f->data_offset = (int64_t)pos;
  f->frame_bytes = (f->has_box ? 56 : 0) + 3 * (8 + 4 * f->num_atoms);
  // the frame count comes from the file size: interrupted writers leave the header stale
  f->num_frames = ((int64_t)f->size - f->data_offset) / f->frame_bytes;
  if (f->num_frames > 0) {
    const unsigned char* frame = p + f->data_offset + (f->has_box ? 56 : 0);
    if (read_i32(frame) != 4 * f->num_atoms) return fail("unexpected coordinate record size");
  }
  *out = f;
  return CVP_OK;
}

int cvp_dcd_info(const cvp_dcd* f, int64_t* num_frames, int64_t* num_atoms, int* has_box,
                 int32_t* first_step, int32_t* interval) {
  CVP_REQUIRE(f != nullptr, "null argument");
  if (num_frames) *num_frames = f->num_frames;
  if (num_atoms) *num_atoms = f->num_atoms;
  if (has_box) *has_box = f->has_box;
  if (first_step) *first_step = f->first_step;
  if (interval) *interval = f->interval;
  return CVP_OK;
}

// Frames first .. first + count - 1 -> positions [count][N][3] float nm (x / 10.0f, correctly
// rounded) and, if the file has unit cells and `cell` is not NULL, the raw cell records
// [count][6] doubles (A, gamma, B, beta, alpha, C as stored).  `threads` <= 0: one per host core.
int cvp_dcd_read(const cvp_dcd* f, int64_t first, int64_t count, float* positions, double* cell,
                 int threads) {
  CVP_REQUIRE(f != nullptr && positions != nullptr, "null argument");
  CVP_REQUIRE(first >= 0 && count >= 0 && first + count <= f->num_frames, "frames outside the file");
  if (count == 0) return CVP_OK;
  int workers = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  workers = (int)std::min<int64_t>(workers, count);
  const int64_t n = f->num_atoms;
  auto convert = [&](int64_t lo, int64_t hi) {
    for (int64_t b = lo; b < hi; ++b) {
      const unsigned char* frame = f->map + f->data_offset + (first + b) * f->frame_bytes;
      if (f->has_box) {
        if (cell) std::memcpy(cell + 6 * b, frame + 4, 48);
        frame += 56;
      }
      const float* x = reinterpret_cast<const float*>(frame + 4);
      const float* y = reinterpret_cast<const float*>(frame + 4 + (8 + 4 * n));
      const float* z = reinterpret_cast<const float*>(frame + 4 + 2 * (8 + 4 * n));
      float* out = positions + b * n * 3;
      for (int64_t i = 0; i < n; ++i) {
        out[3 * i] = x[i] / 10.0f;
        out[3 * i + 1] = y[i] / 10.0f;
        out[3 * i + 2] = z[i] / 10.0f;
      }
    }
  };
  if (workers <= 1) {
    convert(0, count);
    return CVP_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(workers);
  for (int w = 0; w < workers; ++w)
    pool.emplace_back(convert, count * w / workers, count * (w + 1) / workers);
  for (auto& t : pool) t.join();
  return CVP_OK;
}

int cvp_dcd_close(cvp_dcd* f) {
  if (f == nullptr) return CVP_OK;
  if (f->map) munmap(const_cast<unsigned char*>(f->map), f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
  return CVP_OK;
}

}  // extern "C"
