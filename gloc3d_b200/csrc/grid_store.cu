// grid_store.cu -- the grid store file (include/gloc3d.h "grid store file"):
// a map's BEV grids on disk, bit-packed, so that a database is projected once instead of at
// every start-up (the reference re-projects every scan, global_localization.cpp:419-449).
// Host code only; the two store-level functions move the cells through the existing entry
// points (gloc_csm_get_precomputation_grid / gloc_csm_add_grid_u8).
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/gloc3d.h"
#include "common.cuh"

namespace {

constexpr char kMagic[8] = {'G', 'L', 'O', 'C', 'G', 'R', 'D', '1'};
constexpr uint32_t kVersion = 1, kEncRaw = 0, kEncBits = 1;
constexpr int64_t kMaxCells = (int64_t)1 << 31;   // a grid side is an int; refuse absurd headers

struct Record {   // 48 bytes on disk, little-endian
  int32_t nx, ny;
  double resolution, max_x, max_y;
  uint32_t encoding, reserved;
  uint64_t payload_bytes;
};
static_assert(sizeof(Record) == 48, "grid record layout");

bool is_binary(const uint8_t* v, size_t n) {
  for (size_t i = 0; i < n; ++i)
    if (v[i] != 0 && v[i] != 255) return false;
  return true;
}

void pack_bits(const uint8_t* v, size_t n, std::vector<uint8_t>* out) {
  out->assign((n + 7) / 8, 0);
  for (size_t i = 0; i < n; ++i)
    if (v[i]) (*out)[i >> 3] |= (uint8_t)(1u << (i & 7));
}

void unpack_bits(const uint8_t* bits, size_t n, uint8_t* v) {
  for (size_t i = 0; i < n; ++i) v[i] = (bits[i >> 3] >> (i & 7)) & 1u ? 255 : 0;
}

}  // namespace

struct gloc_grid_file {
  FILE* fp = nullptr;
  uint64_t n_grids = 0, next = 0;
  uint32_t tag = 0;
  std::string path;
  std::vector<uint8_t> scratch;
};

using gloc::fail;

extern "C" {

int gloc_grid_file_write(const char* path, const gloc_grid_info* infos, const uint8_t* const* level1,
                         size_t n) {
  return gloc_grid_file_write_tagged(path, infos, level1, n, 0);
}

int gloc_grid_file_write_tagged(const char* path, const gloc_grid_info* infos, const uint8_t* const* level1,
                                size_t n, uint32_t tag) {
  if (!path || (n && (!infos || !level1)))
    return fail(GLOC_ERR_INVALID, "gloc_grid_file_write: null argument");
  for (size_t i = 0; i < n; ++i)
    if (!level1[i] || infos[i].nx < 1 || infos[i].ny < 1 || !(infos[i].resolution > 0))
      return fail(GLOC_ERR_INVALID, "gloc_grid_file_write: grid " + std::to_string(i) + " is empty");
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(GLOC_ERR_INVALID, std::string("gloc_grid_file_write: cannot open ") + path);
  const uint32_t head[2] = {kVersion, tag};
  const uint64_t count = n;
  bool ok = std::fwrite(kMagic, 1, 8, fp) == 8 && std::fwrite(head, 4, 2, fp) == 2 &&
            std::fwrite(&count, 8, 1, fp) == 1;
  std::vector<uint8_t> bits;
  for (size_t i = 0; ok && i < n; ++i) {
    const size_t cells = (size_t)infos[i].nx * (size_t)infos[i].ny;
    Record r;
    r.nx = infos[i].nx;
    r.ny = infos[i].ny;
    r.resolution = infos[i].resolution;
    r.max_x = infos[i].max_x;
    r.max_y = infos[i].max_y;
    r.reserved = 0;
    const uint8_t* payload = level1[i];
    if (is_binary(level1[i], cells)) {
      pack_bits(level1[i], cells, &bits);
      r.encoding = kEncBits;
      r.payload_bytes = bits.size();
      payload = bits.data();
    } else {
      r.encoding = kEncRaw;
      r.payload_bytes = cells;
    }
    ok = std::fwrite(&r, sizeof r, 1, fp) == 1 &&
         std::fwrite(payload, 1, (size_t)r.payload_bytes, fp) == (size_t)r.payload_bytes;
  }
  ok = (std::fclose(fp) == 0) && ok;
  if (!ok) return fail(GLOC_ERR_INVALID, std::string("gloc_grid_file_write: write failed: ") + path);
  return GLOC_OK;
}

int gloc_grid_file_open(const char* path, gloc_grid_file** out, size_t* n_grids) {
  if (!path || !out) return fail(GLOC_ERR_INVALID, "gloc_grid_file_open: null argument");
  *out = nullptr;
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(GLOC_ERR_INVALID, std::string("gloc_grid_file_open: cannot open ") + path);
  char magic[8];
  uint32_t head[2];
  uint64_t count = 0;
  if (std::fread(magic, 1, 8, fp) != 8 || std::memcmp(magic, kMagic, 8) != 0 ||
      std::fread(head, 4, 2, fp) != 2 || std::fread(&count, 8, 1, fp) != 1) {
    std::fclose(fp);
    return fail(GLOC_ERR_INVALID, std::string("gloc_grid_file_open: not a grid store file: ") + path);
  }
  if (head[0] != kVersion) {
    std::fclose(fp);
    return fail(GLOC_ERR_RANGE, "gloc_grid_file_open: unsupported version " + std::to_string(head[0]));
  }
  gloc_grid_file* f = new (std::nothrow) gloc_grid_file;   // the C ABI never throws
  if (!f) {
    std::fclose(fp);
    return fail(GLOC_ERR_NOMEM, "gloc_grid_file_open: out of host memory");
  }
  f->fp = fp;
  f->n_grids = count;
  f->tag = head[1];
  f->path = path;
  *out = f;
  if (n_grids) *n_grids = (size_t)count;
  return GLOC_OK;
}

int gloc_grid_file_next(gloc_grid_file* f, gloc_grid_info* info, uint8_t* level1, size_t capacity) {
  if (!f || !info) return fail(GLOC_ERR_INVALID, "gloc_grid_file_next: null argument");
  if (f->next >= f->n_grids) return fail(GLOC_ERR_RANGE, "gloc_grid_file_next: no more grids");
  const long at = std::ftell(f->fp);
  Record r;
  if (std::fread(&r, sizeof r, 1, f->fp) != 1)
    return fail(GLOC_ERR_INVALID, "gloc_grid_file_next: truncated file: " + f->path);
  const int64_t cells = (int64_t)r.nx * (int64_t)r.ny;
  const uint64_t want = r.encoding == kEncBits ? (uint64_t)((cells + 7) / 8) : (uint64_t)cells;
  if (r.nx < 1 || r.ny < 1 || cells > kMaxCells || !(r.resolution > 0) ||
      (r.encoding != kEncBits && r.encoding != kEncRaw) || r.payload_bytes != want)
    return fail(GLOC_ERR_INVALID, "gloc_grid_file_next: corrupt record " + std::to_string(f->next) + " in " + f->path);
  info->nx = r.nx;
  info->ny = r.ny;
  info->resolution = r.resolution;
  info->max_x = r.max_x;
  info->max_y = r.max_y;
  if (!level1 || capacity < (size_t)cells) {   // sizing call: leave the record unread
    std::fseek(f->fp, at, SEEK_SET);
    return GLOC_OK;
  }
  if (r.encoding == kEncBits) {
    f->scratch.resize((size_t)want);
    if (std::fread(f->scratch.data(), 1, (size_t)want, f->fp) != (size_t)want)
      return fail(GLOC_ERR_INVALID, "gloc_grid_file_next: truncated file: " + f->path);
    unpack_bits(f->scratch.data(), (size_t)cells, level1);
  } else if (std::fread(level1, 1, (size_t)cells, f->fp) != (size_t)cells) {
    return fail(GLOC_ERR_INVALID, "gloc_grid_file_next: truncated file: " + f->path);
  }
  ++f->next;
  return GLOC_OK;
}

uint32_t gloc_grid_file_tag(const gloc_grid_file* f) { return f ? f->tag : 0; }

void gloc_grid_file_close(gloc_grid_file* f) {
  if (!f) return;
  if (f->fp) std::fclose(f->fp);
  delete f;
}

int gloc_csm_save_grids(gloc_csm_store* store, const char* path) {
  return gloc_csm_save_grids_tagged(store, path, 0);
}

int gloc_csm_save_grids_tagged(gloc_csm_store* store, const char* path, uint32_t tag) {
  if (!store || !path) return fail(GLOC_ERR_INVALID, "gloc_csm_save_grids: null argument");
  const int n = gloc_csm_num_grids(store);
  std::vector<gloc_grid_info> infos((size_t)n);
  std::vector<std::vector<uint8_t>> cells((size_t)n);
  std::vector<const uint8_t*> ptrs((size_t)n);
  for (int i = 0; i < n; ++i) {
    int rc = gloc_csm_get_grid_info(store, i, &infos[i]);
    if (rc != GLOC_OK) return rc;
    cells[i].resize((size_t)infos[i].nx * (size_t)infos[i].ny);
    rc = gloc_csm_get_precomputation_grid(store, i, 1, cells[i].data());   // device -> host
    if (rc != GLOC_OK) return rc;
    ptrs[i] = cells[i].data();
  }
  return gloc_grid_file_write_tagged(path, infos.data(), ptrs.data(), (size_t)n, tag);
}

int gloc_csm_load_grids(gloc_csm_store* store, const char* path, int* first_grid_id, int* n_grids) {
  if (!store || !path) return fail(GLOC_ERR_INVALID, "gloc_csm_load_grids: null argument");
  gloc_grid_file* f = nullptr;
  size_t n = 0;
  int rc = gloc_grid_file_open(path, &f, &n);
  if (rc != GLOC_OK) return rc;
  std::vector<uint8_t> cells;
  int first = -1;
  for (size_t i = 0; i < n && rc == GLOC_OK; ++i) {
    gloc_grid_info info;
    rc = gloc_grid_file_next(f, &info, nullptr, 0);
    if (rc != GLOC_OK) break;
    cells.resize((size_t)info.nx * (size_t)info.ny);
    rc = gloc_grid_file_next(f, &info, cells.data(), cells.size());
    if (rc != GLOC_OK) break;
    int gid = -1;
    rc = gloc_csm_add_grid_u8(store, cells.data(), info.nx, info.ny, info.resolution, info.max_x,
                              info.max_y, &gid);   // host -> device
    if (rc == GLOC_OK && first < 0) first = gid;
  }
  gloc_grid_file_close(f);
  if (rc != GLOC_OK) return rc;
  if (first_grid_id) *first_grid_id = first < 0 ? gloc_csm_num_grids(store) : first;
  if (n_grids) *n_grids = (int)n;
  return GLOC_OK;
}

}  // extern "C"
