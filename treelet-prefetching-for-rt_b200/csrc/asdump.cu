// asdump.cu -- host side only: the acceleration-structure dump files of VulkanRayTracing::dump_AS ->
// dump_descriptor_set_for_AS(split_files = true) (vulkan_ray_tracing.cc:4455-4558, offsets from findOffsetBounds :4901-4943):
//   <prefix>.asmain      desc_size bytes starting at the TLAS header
//   <prefix>.asback      bytes [tlas + max_backwards, tlas + min_backwards + back_buffer)   (BLASes below the TLAS)
//   <prefix>.asfront     bytes [tlas + min_forwards,  tlas + max_forwards  + front_buffer)  (BLASes above the TLAS)
//   <prefix>.asmetadata  "desc_size,VkDescriptorType,max_backwards,min_backwards,min_forwards,max_forwards,back_buffer,
//                         front_buffer,haveBackwards,haveForwards"
// max/min_backwards are the most / least negative BLAS offsets from the TLAS, min/max_forwards the smallest / largest
// positive ones (0 = none).  The reader rebuilds one host image with the original relative placement; registration walks the
// TLAS to find the BLAS headers (the dump does not list them).  No sample dumps ship with the reference.
#include "vsrt_internal.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

namespace {
bool write_file(const std::string& path, const void* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb"); if (!f) return false;
  const bool ok = fwrite(p, 1, n, f) == n; fclose(f); return ok;
}
bool read_file(const std::string& path, void* p, size_t n) {
  FILE* f = fopen(path.c_str(), "rb"); if (!f) return false;
  const bool ok = fread(p, 1, n, f) == n && fgetc(f) == EOF; fclose(f); return ok;
}
struct Meta { long long desc_size, type, max_back, min_back, min_fwd, max_fwd, back_buf, front_buf, have_back, have_fwd; };
}  // namespace

extern "C" {

int vsrt_as_dump_write(const char* prefix, const void* tlas, uint64_t desc_size, const void* const* child_addrs, uint32_t n_children,
                       uint64_t back_buffer, uint64_t front_buffer) {
  if (!prefix || !tlas || !desc_size || (n_children && !child_addrs)) return VSRT_E_INVALID;
  // findOffsetBounds (:4901-4943)
  std::vector<long long> pos, neg;
  for (uint32_t i = 0; i < n_children; i++) { const long long off = (long long)((uint64_t)(uintptr_t)child_addrs[i] - (uint64_t)(uintptr_t)tlas); (off >= 0 ? pos : neg).push_back(off); }
  std::sort(pos.begin(), pos.end()); std::sort(neg.begin(), neg.end());
  Meta m;
  m.desc_size = (long long)desc_size; m.type = 1000150000;                           // VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_KHR
  m.max_back = neg.empty() ? 0 : neg.front(); m.min_back = neg.empty() ? 0 : neg.back();
  m.min_fwd = pos.empty() ? 0 : pos.front(); m.max_fwd = pos.empty() ? 0 : pos.back();
  m.back_buf = (long long)back_buffer; m.front_buf = (long long)front_buffer;
  m.have_back = (m.max_back != 0) && (m.min_back != 0); m.have_fwd = (m.min_fwd != 0) && (m.max_fwd != 0);   // :4487-4488
  const std::string p(prefix);
  const uint8_t* t = (const uint8_t*)tlas;
  if (!write_file(p + ".asmain", t, desc_size)) return VSRT_E_INVALID;
  if (m.have_back && !write_file(p + ".asback", t + m.max_back, (size_t)(m.min_back - m.max_back + m.back_buf))) return VSRT_E_INVALID;
  if (m.have_fwd && !write_file(p + ".asfront", t + m.min_fwd, (size_t)(m.max_fwd - m.min_fwd + m.front_buf))) return VSRT_E_INVALID;
  char line[256];
  const int n = snprintf(line, sizeof(line), "%d,%d,%ld,%ld,%ld,%ld,%ld,%ld,%d,%d", (int)m.desc_size, (int)m.type, (long)m.max_back, (long)m.min_back,
                         (long)m.min_fwd, (long)m.max_fwd, (long)m.back_buf, (long)m.front_buf, (int)m.have_back, (int)m.have_fwd);     // :4521-4531
  return write_file(p + ".asmetadata", line, (size_t)n) ? VSRT_OK : VSRT_E_INVALID;
}

int vsrt_as_dump_read(const char* prefix, void** image, uint64_t* image_size, uint64_t* tlas_offset) {
  if (!prefix || !image || !image_size || !tlas_offset) return VSRT_E_INVALID;
  *image = nullptr; *image_size = 0; *tlas_offset = 0;
  const std::string p(prefix);
  FILE* f = fopen((p + ".asmetadata").c_str(), "r"); if (!f) return VSRT_E_INVALID;
  Meta m;
  const int got = fscanf(f, "%lld,%lld,%lld,%lld,%lld,%lld,%lld,%lld,%lld,%lld", &m.desc_size, &m.type, &m.max_back, &m.min_back, &m.min_fwd, &m.max_fwd,
                         &m.back_buf, &m.front_buf, &m.have_back, &m.have_fwd);
  fclose(f);
  if (got != 10 || m.desc_size < 64 || m.max_back > 0 || m.min_back > 0 || m.min_fwd < 0 || m.max_fwd < m.min_fwd || m.max_back > m.min_back ||
      m.back_buf < 0 || m.front_buf < 0) return VSRT_E_INVALID;
  const long long lo = m.have_back ? m.max_back : 0;
  long long hi = m.desc_size;
  if (m.have_back) hi = std::max(hi, m.min_back + m.back_buf);
  if (m.have_fwd) hi = std::max(hi, m.max_fwd + m.front_buf);
  if ((lo & 63) != 0) return VSRT_E_INVALID;                                         // BLAS headers are 64-byte aligned
  const uint64_t size = ((uint64_t)(hi - lo) + 63) & ~63ull;
  uint8_t* img = (uint8_t*)aligned_alloc(64, size); if (!img) return VSRT_E_INVALID;
  memset(img, 0, size);
  uint8_t* t = img - lo;
  bool ok = read_file(p + ".asmain", t, (size_t)m.desc_size);
  if (ok && m.have_back) ok = read_file(p + ".asback", t + m.max_back, (size_t)(m.min_back - m.max_back + m.back_buf));
  if (ok && m.have_fwd) ok = read_file(p + ".asfront", t + m.min_fwd, (size_t)(m.max_fwd - m.min_fwd + m.front_buf));
  if (!ok) { free(img); return VSRT_E_INVALID; }
  *image = img; *image_size = size; *tlas_offset = (uint64_t)(-lo);
  return VSRT_OK;
}

void vsrt_as_dump_free(void* image) { free(image); }

// Registers the TLAS at image + tlas_offset and every BLAS an instance leaf of it references (found by walking the TLAS
// like createTreelets does, :886-992).  Each registration runs to the end of the image: the dump has no per-BLAS sizes.
int vsrt_register_as_image(vsrt_context* ctx, const void* image, uint64_t image_size, uint64_t tlas_offset, int64_t device_delta, uint32_t* n_blas_out) {
  if (!ctx || !image || tlas_offset + 64 > image_size || (tlas_offset & 63)) return VSRT_E_INVALID;
  const uint8_t* img = (const uint8_t*)image;
  const uint8_t* tlas = img + tlas_offset;
  uint64_t root_off; memcpy(&root_off, tlas, 8);
  if ((root_off & 63) || tlas_offset + root_off + 64 > image_size) return VSRT_E_BAD_BVH;
  std::vector<uint64_t> stack{ tlas_offset + root_off }, blas;
  size_t visited = 0;
  while (!stack.empty()) {
    const uint64_t off = stack.back(); stack.pop_back();
    if (++visited > image_size / 64) return VSRT_E_BAD_BVH;                          // a cycle
    const uint8_t* n = img + off;
    int32_t child_offset; memcpy(&child_offset, n + 12, 4);
    int64_t child = (int64_t)off + (int64_t)child_offset * 64;
    for (int i = 0; i < 6; i++) {
      const uint32_t info = n[22 + i] & 0x3fu, sz = info & 3u, ty = info >> 2;
      if (!sz) continue;
      if (child < 0 || (uint64_t)child + 64ull * sz > image_size) return VSRT_E_BAD_BVH;
      if (ty == 0) stack.push_back((uint64_t)child);
      else if (ty == 1) {
        uint64_t rel; memcpy(&rel, img + child + 64, 8);                              // BVHAddress, relative to the leaf (:1902)
        const uint64_t hdr = (uint64_t)child + rel;
        if (rel == 0 || (hdr & 63) || hdr + 64 > image_size) return VSRT_E_BAD_BVH;
        blas.push_back(hdr);
      } else return VSRT_E_BAD_BVH;                                                   // a TLAS leaf must be an instance (:926)
      child += 64ll * sz;
    }
  }
  std::sort(blas.begin(), blas.end()); blas.erase(std::unique(blas.begin(), blas.end()), blas.end());
  int rc = vsrt_alloc_tlas(ctx, tlas, image_size - tlas_offset, (uint64_t)(uintptr_t)tlas + (uint64_t)device_delta);
  for (size_t i = 0; i < blas.size() && rc == VSRT_OK; i++)
    rc = vsrt_alloc_blas(ctx, img + blas[i], image_size - blas[i], (uint64_t)(uintptr_t)(img + blas[i]) + (uint64_t)device_delta);
  if (n_blas_out) *n_blas_out = (uint32_t)blas.size();
  return rc;
}

}  // extern "C"
