// Native safetensors reader: mmap the checkpoint, parse its JSON header, and copy tensors
// straight from the mapping into caller-provided DEVICE memory — no intermediate host tensors, no Python-side copies.
// Replaces the file-reading half of the reference's loader (moondream/torch/weights.py:156-171, which goes through the
// `safetensors` package); key-layout normalisation (canonical / legacy HF names, weights.py:36-131) stays in Python
// (moondream_b200/weights.py), it is a table lookup.
//
// File format (safetensors): u64 little-endian header length N, N bytes of JSON
//   {"name": {"dtype": "BF16", "shape": [..], "data_offsets": [begin, end]}, ..., "__metadata__": {...}}
// followed by the tensor bytes; offsets are relative to the end of the header.
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

#include "../../include/moondream_b200.h"
#include "kernels.cuh"

namespace md {

struct StEntry {
  std::string name, dtype;
  std::vector<long long> shape;
  long long begin = 0, end = 0;
};

struct StFile {
  int fd = -1;
  const unsigned char* map = nullptr;
  size_t size = 0;
  size_t data0 = 0;                       // file offset of the tensor bytes
  std::vector<StEntry> entries;
};

// ---- a JSON reader for exactly what a safetensors header contains ----
struct Json {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  std::string str() {
    std::string s;
    ws();
    if (p >= e || *p != '"') { ok = false; return s; }
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        ++p;
        switch (*p) {
          case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
          case 'b': s += '\b'; break; case 'f': s += '\f'; break;
          case 'u': s += '?'; p += (p + 4 < e ? 4 : 0); break;        // names here are ASCII; keep the length right
          default: s += *p;
        }
        ++p;
      } else {
        s += *p++;
      }
    }
    if (p >= e) { ok = false; return s; }
    ++p;
    return s;
  }
  long long num() {
    ws();
    long long v = 0;
    bool neg = false, any = false;
    if (p < e && *p == '-') { neg = true; ++p; }
    while (p < e && *p >= '0' && *p <= '9') {
      if (v > (0x7fffffffffffffffLL - 9) / 10) { ok = false; return 0; }      // would overflow: not a size any file has
      v = v * 10 + (*p - '0'); ++p; any = true;
    }
    if (!any) ok = false;
    return neg ? -v : v;
  }
  int depth = 0;
  void skip() {                           // any value (used for __metadata__); nesting is bounded, the file is untrusted
    ws();
    if (p >= e || depth > 64) { ok = false; return; }
    if (*p == '"') { str(); return; }
    if (*p == '{' || *p == '[') {
      const char open = *p, close = open == '{' ? '}' : ']';
      ++p;
      ws();
      if (p < e && *p == close) { ++p; return; }
      ++depth;
      while (ok) {
        if (open == '{') { str(); if (!eat(':')) { ok = false; break; } }
        skip();
        if (!ok) break;
        if (eat(',')) continue;
        if (eat(close)) { --depth; return; }
        ok = false;
      }
      --depth;
      return;
    }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;      // number / true / false / null
  }
};

static int parse_header(StFile& f, const char* js, size_t n) {
  Json j{js, js + n};
  if (!j.eat('{')) return set_error("safetensors: header is not a JSON object");
  if (j.eat('}')) return 0;
  while (j.ok) {
    const std::string name = j.str();
    if (!j.eat(':')) return set_error("safetensors: malformed header");
    if (name == "__metadata__") {
      j.skip();
    } else {
      StEntry en;
      en.name = name;
      int have = 0;                         // 1 dtype, 2 shape, 4 data_offsets: all three are required
      if (!j.eat('{')) return set_error("safetensors: tensor entry is not an object");
      while (j.ok) {
        const std::string key = j.str();
        if (!j.eat(':')) return set_error("safetensors: malformed tensor entry");
        if (key == "dtype") { en.dtype = j.str(); have |= 1; }
        else if (key == "shape") {
          have |= 2;
          if (!j.eat('[')) return set_error("safetensors: shape is not an array");
          if (!j.eat(']')) {
            do { en.shape.push_back(j.num()); } while (j.eat(','));
            if (!j.eat(']')) return set_error("safetensors: malformed shape");
          }
        } else if (key == "data_offsets") {
          have |= 4;
          if (!j.eat('[')) return set_error("safetensors: data_offsets is not an array");
          en.begin = j.num();
          if (!j.eat(',')) return set_error("safetensors: malformed data_offsets");
          en.end = j.num();
          if (!j.eat(']')) return set_error("safetensors: malformed data_offsets");
        } else {
          j.skip();
        }
        if (j.eat(',')) continue;
        if (j.eat('}')) break;
        return set_error("safetensors: malformed tensor entry");
      }
      if (!j.ok) break;
      if (have != 7) return set_error("safetensors: a tensor entry lacks dtype, shape or data_offsets");
      f.entries.push_back(en);
    }
    if (j.eat(',')) continue;
    if (j.eat('}')) break;
    return set_error("safetensors: malformed header");
  }
  j.ws();
  while (j.p < j.e && (*j.p == '\0' || *j.p == ' ')) ++j.p;        // writers pad the header to 8 bytes (spaces; some NULs)
  if (j.ok && j.p != j.e) return set_error("safetensors: bytes after the header object");
  return j.ok ? 0 : set_error("safetensors: malformed header");
}

static int dtype_size(const std::string& d) {
  if (d == "BF16" || d == "F16" || d == "I16" || d == "U16") return 2;
  if (d == "F32" || d == "I32" || d == "U32") return 4;
  if (d == "F64" || d == "I64" || d == "U64") return 8;
  if (d == "I8" || d == "U8" || d == "BOOL" || d == "F8_E4M3" || d == "F8_E5M2") return 1;
  return 0;
}

}  // namespace md

struct md_file : md::StFile {};

extern "C" {

int md_safetensors_open(const char* path, md_file** out) {
  if (!path || !out) return md::set_error("md_safetensors_open: null pointer");
  md_file* f = new (std::nothrow) md_file();
  if (!f) return md::set_error("md_safetensors_open: out of host memory");
  f->fd = open(path, O_RDONLY);
  struct stat st;
  if (f->fd < 0 || fstat(f->fd, &st) != 0 || st.st_size < 8) {
    if (f->fd >= 0) close(f->fd);
    delete f;
    return md::set_error("md_safetensors_open: cannot open the file (or it is shorter than a header)");
  }
  f->size = static_cast<size_t>(st.st_size);
  void* m = mmap(nullptr, f->size, PROT_READ, MAP_PRIVATE, f->fd, 0);
  if (m == MAP_FAILED) { close(f->fd); delete f; return md::set_error("md_safetensors_open: mmap failed"); }
  f->map = static_cast<const unsigned char*>(m);
  unsigned long long n = 0;
  for (int i = 7; i >= 0; --i) n = (n << 8) | f->map[i];
  int rc = 0;
  if (n > f->size - 8) rc = md::set_error("md_safetensors_open: header length exceeds the file");
  if (!rc) {
    f->data0 = 8 + static_cast<size_t>(n);
    rc = md::parse_header(*f, reinterpret_cast<const char*>(f->map + 8), static_cast<size_t>(n));
  }
  for (const md::StEntry& en : f->entries) {
    if (rc) break;
    long long elems = 1;
    bool bad = false;
    for (long long d : en.shape) bad = bad || d < 0 || __builtin_mul_overflow(elems, d, &elems);
    const int sz = md::dtype_size(en.dtype);
    long long bytes = 0;
    if (sz) bad = bad || __builtin_mul_overflow(elems, static_cast<long long>(sz), &bytes);
    if (bad || en.begin < 0 || en.end < en.begin || static_cast<unsigned long long>(en.end) > f->size - f->data0 ||
        (sz && bytes != en.end - en.begin))
      rc = md::set_error("md_safetensors_open: a tensor's offsets do not fit its shape or the file");
  }
  if (rc) { munmap(const_cast<unsigned char*>(f->map), f->size); close(f->fd); delete f; return rc; }
  madvise(const_cast<unsigned char*>(f->map), f->size, MADV_SEQUENTIAL);
  *out = f;
  return 0;
}

int md_safetensors_count(const md_file* f) { return f ? static_cast<int>(f->entries.size()) : -1; }

int md_safetensors_info(const md_file* f, int index, const char** name, const char** dtype, int* ndim,
                        long long* shape8, long long* nbytes) {
  if (!f || index < 0 || index >= static_cast<int>(f->entries.size())) return md::set_error("md_safetensors_info: bad index");
  const md::StEntry& en = f->entries[index];
  if (en.shape.size() > 8) return md::set_error("md_safetensors_info: more than 8 dimensions");
  if (name) *name = en.name.c_str();
  if (dtype) *dtype = en.dtype.c_str();
  if (ndim) *ndim = static_cast<int>(en.shape.size());
  if (shape8) for (size_t i = 0; i < en.shape.size(); ++i) shape8[i] = en.shape[i];
  if (nbytes) *nbytes = en.end - en.begin;
  return 0;
}

int md_safetensors_find(const md_file* f, const char* name) {
  if (!f || !name) return -1;
  for (size_t i = 0; i < f->entries.size(); ++i)
    if (f->entries[i].name == name) return static_cast<int>(i);
  return -1;
}

int md_safetensors_read(const md_file* f, int index, void* dst, long long dst_bytes, int dst_is_device, void* stream) {
  if (!f || !dst || index < 0 || index >= static_cast<int>(f->entries.size())) return md::set_error("md_safetensors_read: bad argument");
  const md::StEntry& en = f->entries[index];
  const long long n = en.end - en.begin;
  if (dst_bytes != n) return md::set_error("md_safetensors_read: destination size differs from the tensor's");
  const unsigned char* src = f->map + f->data0 + en.begin;
  if (n == 0) return 0;
  if (!dst_is_device) { memcpy(dst, src, static_cast<size_t>(n)); return 0; }
  cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(n), cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return md::set_error(cudaGetErrorString(e));
  return 0;
}

void md_safetensors_close(md_file* f) {
  if (!f) return;
  if (f->map) munmap(const_cast<unsigned char*>(f->map), f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
}

}  // extern "C"
