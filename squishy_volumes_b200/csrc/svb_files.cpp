// svb_files.cpp — frame files, input files, scene set-up and keyframes (include/svb_files.h), host only.
//
// Written from the reference's serde data model and the published bincode 1.3.3 wire format (little endian, fixed-width
// integers, u64 lengths, u8 Option tags, u32 enum variant indices); every routine cites the Rust item it follows.
#include "../../include/svb_files.h"

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace {

constexpr size_t MAGIC_LEN = 32, VERSION_LEN = 64, DATA_OFFSET = MAGIC_LEN + VERSION_LEN;  // file_util/src/lib.rs:26-28
const char FRAME_MAGIC[MAGIC_LEN + 1] = "Squishy Volumes Frame File Magic";               // file_frame/src/io_state.rs:20-29
const char INPUT_MAGIC[MAGIC_LEN + 1] = "Squishy Volumes Input File Magic";               // file_input/src/common.rs
const char DEFAULT_VERSION[] = "0.3.4";                                                   // file_util/Cargo.toml:3

// file_frame/src/particles.rs:23-33
constexpr uint32_t F_IS_SOLID = 1, F_IS_FLUID = 2, F_USE_VISCOSITY = 4, F_USE_SAND_ALPHA = 8, F_ALL = 0x7f;

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
  char buf[768];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

// ---------------------------------------------------------------------------------------------- bincode cursor / sink
struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  Cursor(const uint8_t* b, size_t n) : p(b), end(b + n) {}
  bool take(void* dst, size_t n) {
    if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
    if (n) std::memcpy(dst, p, n);
    p += n;
    return true;
  }
  template <class T> T get() { T v{}; take(&v, sizeof(T)); return v; }
  // Vec<[T; K]>: u64 length, then the elements back to back.  A length that cannot fit the remaining bytes is malformed.
  template <class T> bool vec(std::vector<T>& out, size_t per_element) {
    const uint64_t n = get<uint64_t>();
    if (!ok || per_element == 0 || n > (uint64_t)(end - p) / (per_element * sizeof(T))) { ok = false; return false; }
    out.resize((size_t)n * per_element);
    return take(out.data(), out.size() * sizeof(T));
  }
  // the same Vec, left where it is: element count and a pointer into the buffer
  bool vec_view(size_t element_bytes, uint64_t& count, const uint8_t*& data) {
    count = get<uint64_t>();
    if (!ok || element_bytes == 0 || count > (uint64_t)(end - p) / element_bytes) { ok = false; count = 0; data = nullptr; return false; }
    data = p;
    p += (size_t)count * element_bytes;
    return true;
  }
  template <class T> bool opt_vec(bool& present, std::vector<T>& out, size_t per_element) {
    const uint8_t tag = get<uint8_t>();
    if (tag > 1) ok = false;
    present = tag == 1;
    if (!present) { out.clear(); return ok; }
    return vec(out, per_element);
  }
  bool string(std::string& s) {
    const uint64_t n = get<uint64_t>();
    if (!ok || n > (uint64_t)(end - p)) { ok = false; return false; }
    s.assign(reinterpret_cast<const char*>(p), (size_t)n);
    p += n;
    return true;
  }
};

struct Sink {
  FILE* f = nullptr;
  bool ok = true;
  uint64_t pos = 0;
  void put(const void* src, size_t n) {
    if (!ok || !n) return;
    if (fwrite(src, 1, n, f) != n) ok = false;
    pos += n;
  }
  template <class T> void val(T v) { put(&v, sizeof(T)); }
  template <class T> void vec(const T* data, uint64_t n, size_t per_element) {
    val<uint64_t>(n);
    if (data) put(data, (size_t)n * per_element * sizeof(T));
    else {  // a missing optional array is written as zeros
      std::vector<T> z(4096 * per_element, T{});
      for (uint64_t done = 0; done < n;) {
        const uint64_t c = std::min<uint64_t>(4096, n - done);
        put(z.data(), (size_t)c * per_element * sizeof(T));
        done += c;
      }
    }
  }
  template <class T> void opt_vec(const T* data, uint64_t n, size_t per_element) {
    val<uint8_t>(data ? 1 : 0);
    if (data) vec(data, n, per_element);
  }
  void string(const std::string& s) {
    val<uint64_t>(s.size());
    put(s.data(), s.size());
  }
};

// file_util/src/lib.rs:38-43, 65-74
bool version_bytes(const char* version, uint8_t out[VERSION_LEN]) {
  const char* v = version ? version : DEFAULT_VERSION;
  const size_t n = std::strlen(v);
  if (n > VERSION_LEN) return false;
  std::memset(out, 0, VERSION_LEN);
  std::memcpy(out, v, n);
  return true;
}
// file_util/src/lib.rs:76-101
int check_magic_and_version(Cursor& c, const char* magic, const char* version, const char* path) {
  uint8_t m[MAGIC_LEN], v[VERSION_LEN], want[VERSION_LEN];
  if (!c.take(m, MAGIC_LEN)) return fail(SVBF_IO_ERROR, "%s: failed to read magic bytes", path);
  if (std::memcmp(m, magic, MAGIC_LEN) != 0) return fail(SVBF_MAGIC_MISMATCH, "%s: magic bytes mismatch, expected \"%s\"", path, magic);
  if (!c.take(v, VERSION_LEN)) return fail(SVBF_IO_ERROR, "%s: failed to read version bytes", path);
  if (!version_bytes(version, want)) return fail(SVBF_BAD_ARGUMENT, "version string too long");
  if (std::memcmp(v, want, VERSION_LEN) != 0) {
    std::string found(reinterpret_cast<const char*>(v), strnlen(reinterpret_cast<const char*>(v), VERSION_LEN));
    return fail(SVBF_VERSION_MISMATCH, "%s: version mismatch, found %s, but expected %s", path, found.c_str(), version ? version : DEFAULT_VERSION);
  }
  return 0;
}

int read_range(FILE* f, uint64_t offset, uint64_t n, std::vector<uint8_t>& out, const char* path) {
  out.resize((size_t)n);
  if (fseeko(f, (off_t)offset, SEEK_SET) != 0) return fail(SVBF_IO_ERROR, "%s: seek to %llu failed: %s", path, (unsigned long long)offset, std::strerror(errno));
  if (n && fread(out.data(), 1, (size_t)n, f) != (size_t)n) return fail(SVBF_IO_ERROR, "%s: short read at %llu", path, (unsigned long long)offset);
  return 0;
}

struct FileCloser {
  void operator()(FILE* f) const { if (f) fclose(f); }
};
using FilePtr = std::unique_ptr<FILE, FileCloser>;

// ---------------------------------------------------------------------------------------------- input data model
struct ParticlesInputData {  // file_input/src/frame.rs:10-26
  std::vector<uint32_t> flags;
  bool has_transforms = false, has_sizes = false, has_densities = false, has_youngs = false, has_poissons = false, has_initial_positions = false, has_initial_velocities = false,
       has_visc_dynamic = false, has_visc_bulk = false, has_exponents = false, has_bulk = false, has_sand = false, has_goals = false;
  std::vector<float> transforms, sizes, densities, youngs, poissons, initial_positions, initial_velocities, visc_dynamic, visc_bulk, bulk, sand, goals;
  std::vector<uint32_t> exponents;
};
struct ColliderInputData {  // file_input/src/collider_inputs.rs:12-18
  std::vector<float> vertex_positions, frictions, dampings;
  std::vector<uint32_t> triangle_indices;
};
struct InputFrameData {  // file_input/src/frame.rs:44-49
  float gravity[3] = {0, 0, 0};
  std::map<std::string, ParticlesInputData> particles;
  std::map<std::string, ColliderInputData> colliders;
};
struct ObjectInfo {  // InputObject + InputRange (file_input/src/header.rs:74-116)
  int kind = 0;
  uint64_t count = 0, count2 = 0, start = 0, start2 = 0;
};

bool parse_particles_input(Cursor& c, ParticlesInputData& d) {
  c.vec(d.flags, 1);
  c.opt_vec(d.has_transforms, d.transforms, 16);
  c.opt_vec(d.has_sizes, d.sizes, 1);
  c.opt_vec(d.has_densities, d.densities, 1);
  c.opt_vec(d.has_youngs, d.youngs, 1);
  c.opt_vec(d.has_poissons, d.poissons, 1);
  c.opt_vec(d.has_initial_positions, d.initial_positions, 3);
  c.opt_vec(d.has_initial_velocities, d.initial_velocities, 3);
  c.opt_vec(d.has_visc_dynamic, d.visc_dynamic, 1);
  c.opt_vec(d.has_visc_bulk, d.visc_bulk, 1);
  c.opt_vec(d.has_exponents, d.exponents, 1);
  c.opt_vec(d.has_bulk, d.bulk, 1);
  c.opt_vec(d.has_sand, d.sand, 1);
  c.opt_vec(d.has_goals, d.goals, 3);
  return c.ok;
}
bool parse_frame(Cursor& c, InputFrameData& fr) {
  c.take(fr.gravity, 12);
  const uint64_t np = c.get<uint64_t>();
  for (uint64_t i = 0; c.ok && i < np; ++i) {
    std::string name;
    c.string(name);
    parse_particles_input(c, fr.particles[name]);
  }
  const uint64_t nc = c.get<uint64_t>();
  for (uint64_t i = 0; c.ok && i < nc; ++i) {
    std::string name;
    c.string(name);
    ColliderInputData& d = fr.colliders[name];
    c.vec(d.vertex_positions, 3);
    c.vec(d.triangle_indices, 3);
    c.vec(d.frictions, 1);
    c.vec(d.dampings, 1);
  }
  return c.ok;
}

}  // namespace

struct SvbfFrame {
  double time = 0;
  std::vector<uint8_t> file;   // the whole frame file: the bulk arrays are used in place, only the parameters are unpacked
  struct View { uint64_t count = 0; const uint8_t* data = nullptr; };
  View flags, energies, bits, x, F, v, C, x0, node_ids, node_bits, node_masses, node_velocities;
  std::vector<float> mass, volume, p0, p1, alpha, vd, vb;
  bool has_grid = false;
};

struct SvbfInput {
  std::string path;
  FilePtr file;
  uint64_t size = 0, index_offset = 0;
  std::vector<uint64_t> offsets;
  SvbConsts consts{};
  std::map<std::string, ObjectInfo> objects;
  uint64_t total_particles = 0, total_vertices = 0, total_triangles = 0;
  std::string version;
  bool has_version = false;

  int read_frame(uint64_t frame, InputFrameData& out) {  // InputReader::read_frame, reading.rs:60-70
    if (frame >= offsets.size()) return fail(SVBF_FRAME_NOT_AVAILABLE, "%s: frame %llu requested, %zu available", path.c_str(), (unsigned long long)frame, offsets.size());
    const uint64_t begin = offsets[(size_t)frame];
    const uint64_t stop = frame + 1 < offsets.size() ? offsets[(size_t)frame + 1] : index_offset;
    if (begin > stop || stop > size) return fail(SVBF_FORMAT, "%s: frame %llu has a bad offset", path.c_str(), (unsigned long long)frame);
    std::vector<uint8_t> buf;
    if (int rc = read_range(file.get(), begin, stop - begin, buf, path.c_str())) return rc;
    Cursor c(buf.data(), buf.size());
    if (!parse_frame(c, out)) return fail(SVBF_FORMAT, "%s: frame %llu is truncated or malformed", path.c_str(), (unsigned long long)frame);
    return 0;
  }
};

struct SvbfInputWriter {
  std::string path;
  Sink sink;
  FilePtr file;
  std::map<std::string, ObjectInfo> objects;
  std::vector<uint64_t> offsets;
};

extern "C" {

const char* svbf_last_error(void) { return g_error.c_str(); }
const char* svbf_default_version(void) { return DEFAULT_VERSION; }

// ================================================================================================ frame files
int32_t svbf_frame_path(const char* cache_dir, uint64_t frame, char* out, size_t cap) {
  if (!cache_dir || !out) return fail(SVBF_BAD_ARGUMENT, "svbf_frame_path: null argument");
  const int n = snprintf(out, cap, "%s/frame_%05llu.bin", cache_dir, (unsigned long long)frame);  // cache/src/util.rs:9-11
  return n < 0 || (size_t)n >= cap ? fail(SVBF_BAD_ARGUMENT, "svbf_frame_path: buffer too small") : 0;
}

int32_t svbf_frame_write(const char* path, const char* version, double time, const SvbParticles* p, const SvbGrid* grid, uint64_t* written_bytes) {
  if (!path || !p) return fail(SVBF_BAD_ARGUMENT, "svbf_frame_write: null argument");
  uint8_t ver[VERSION_LEN];
  if (!version_bytes(version, ver)) return fail(SVBF_BAD_ARGUMENT, "version string too long");
  const uint64_t n = p->n;
  if (n && (!p->flags || !p->mass || !p->initial_volume || !p->mu_or_bulk_modulus || !p->lambda_or_exponent || !p->positions || !p->position_gradients || !p->velocities ||
            !p->velocity_gradients))
    return fail(SVBF_BAD_ARGUMENT, "svbf_frame_write: a required particle array is NULL");
  // io_state.rs:33-38: the frame is written next to its final place and renamed, so readers never see a partial file
  std::string dir(path);
  const size_t slash = dir.find_last_of('/');
  if (slash == std::string::npos) dir = ".";
  else dir.resize(slash ? slash : 1);
  const std::string temp = dir + "/temp.bin";
  Sink s;
  s.f = fopen(temp.c_str(), "wb");
  if (!s.f) return fail(SVBF_IO_ERROR, "failed to create %s: %s", temp.c_str(), std::strerror(errno));
  FilePtr guard(s.f);
  std::vector<char> big(1 << 20);
  setvbuf(s.f, big.data(), _IOFBF, big.size());
  s.put(FRAME_MAGIC, MAGIC_LEN);
  s.put(ver, VERSION_LEN);
  s.val<double>(time);
  // Particles (file_frame/src/particles.rs:93-109), field by field
  s.vec(p->flags, n, 1);
  s.val<uint64_t>(n);  // parameters: Vec<ParticleParameters> (particles.rs:62-84), variable length per particle
  {
    // at most 29 bytes per particle: fill a fixed buffer through a raw pointer (this loop is the serial part of a frame write)
    constexpr size_t CHUNK = 1 << 20, MAX_RECORD = 32;
    std::vector<uint8_t> chunk(CHUNK + MAX_RECORD);
    uint8_t* w = chunk.data();
    auto push4 = [&w](const void* src) { std::memcpy(w, src, 4); w += 4; };
    const uint32_t solid = 0, fluid = 1;
    for (uint64_t i = 0; i < n; ++i) {
      const uint32_t fl = p->flags[i];
      push4(&p->mass[i]);
      push4(&p->initial_volume[i]);
      const bool visc = (fl & F_USE_VISCOSITY) != 0;  // Option<ViscosityParameters>
      *w++ = visc ? 1 : 0;
      if (visc) {
        const float d = p->viscosity_dynamic ? p->viscosity_dynamic[i] : 0.f, b = p->viscosity_bulk ? p->viscosity_bulk[i] : 0.f;
        push4(&d);
        push4(&b);
      }
      if (fl & F_IS_FLUID) {  // SpecificParticleParameters::Fluid { exponent: i32, bulk_modulus: f32 } = variant 1
        const int32_t exponent = (int32_t)p->lambda_or_exponent[i];
        push4(&fluid);
        push4(&exponent);
        push4(&p->mu_or_bulk_modulus[i]);
      } else {  // Solid { mu, lambda, sand_alpha: Option<f32> } = variant 0 (also the Default)
        push4(&solid);
        push4(&p->mu_or_bulk_modulus[i]);
        push4(&p->lambda_or_exponent[i]);
        const bool sand = (fl & F_USE_SAND_ALPHA) != 0;
        *w++ = sand ? 1 : 0;
        if (sand) {
          const float a = p->sand_alpha ? p->sand_alpha[i] : 0.f;
          push4(&a);
        }
      }
      if ((size_t)(w - chunk.data()) >= CHUNK) { s.put(chunk.data(), (size_t)(w - chunk.data())); w = chunk.data(); }
    }
    s.put(chunk.data(), (size_t)(w - chunk.data()));
  }
  s.vec(p->elastic_energies, n, 1);
  s.vec(p->collider_bits, n, 1);
  s.vec(p->positions, n, 3);
  s.vec(p->position_gradients, n, 9);
  s.vec(p->velocities, n, 3);
  s.vec(p->velocity_gradients, n, 9);
  s.vec(p->initial_positions, n, 3);
  // grid_nodes: Option<GridNodes> (file_frame/src/grid_nodes.rs:9-15)
  s.val<uint8_t>(grid ? 1 : 0);
  if (grid) {
    s.vec(grid->node_ids, grid->n, 3);
    s.vec(grid->collider_bits, grid->n, 1);
    s.vec(grid->masses, grid->n, 1);
    s.vec(grid->velocities, grid->n, 3);
  }
  const bool flushed = fflush(s.f) == 0;
  guard.reset();
  if (!s.ok || !flushed) {
    remove(temp.c_str());
    return fail(SVBF_IO_ERROR, "failed to write %s", temp.c_str());
  }
  if (rename(temp.c_str(), path) != 0) return fail(SVBF_IO_ERROR, "failed to move %s to %s: %s", temp.c_str(), path, std::strerror(errno));
  if (written_bytes) *written_bytes = s.pos;
  return 0;
}

int32_t svbf_frame_open(const char* path, const char* version, SvbfFrame** out) {
  if (!path || !out) return fail(SVBF_BAD_ARGUMENT, "svbf_frame_open: null argument");
  *out = nullptr;
  FilePtr f(fopen(path, "rb"));
  if (!f) return fail(SVBF_IO_ERROR, "failed to open %s: %s", path, std::strerror(errno));
  fseeko(f.get(), 0, SEEK_END);
  const uint64_t size = (uint64_t)ftello(f.get());
  std::unique_ptr<SvbfFrame> fr(new SvbfFrame());
  if (int rc = read_range(f.get(), 0, size, fr->file, path)) return rc;
  Cursor c(fr->file.data(), fr->file.size());
  if (int rc = check_magic_and_version(c, FRAME_MAGIC, version, path)) return rc;
  fr->time = c.get<double>();
  c.vec_view(4, fr->flags.count, fr->flags.data);
  const uint64_t n = c.get<uint64_t>();
  if (!c.ok || n > (uint64_t)(c.end - c.p) / 17) return fail(SVBF_FORMAT, "%s: truncated or malformed particle parameters", path);
  fr->mass.resize(n); fr->volume.resize(n); fr->p0.assign(n, 0.f); fr->p1.assign(n, 0.f); fr->alpha.assign(n, 0.f); fr->vd.assign(n, 0.f); fr->vb.assign(n, 0.f);
  for (uint64_t i = 0; c.ok && i < n; ++i) {
    fr->mass[i] = c.get<float>();
    fr->volume[i] = c.get<float>();
    const uint8_t visc = c.get<uint8_t>();
    if (visc > 1) { c.ok = false; break; }
    if (visc) { fr->vd[i] = c.get<float>(); fr->vb[i] = c.get<float>(); }
    const uint32_t variant = c.get<uint32_t>();
    if (variant == 0) {
      fr->p0[i] = c.get<float>();
      fr->p1[i] = c.get<float>();
      const uint8_t sand = c.get<uint8_t>();
      if (sand > 1) { c.ok = false; break; }
      if (sand) fr->alpha[i] = c.get<float>();
    } else if (variant == 1) {
      fr->p1[i] = (float)c.get<int32_t>();
      fr->p0[i] = c.get<float>();
    } else c.ok = false;
  }
  c.vec_view(4, fr->energies.count, fr->energies.data);
  c.vec_view(4, fr->bits.count, fr->bits.data);
  c.vec_view(12, fr->x.count, fr->x.data);
  c.vec_view(36, fr->F.count, fr->F.data);
  c.vec_view(12, fr->v.count, fr->v.data);
  c.vec_view(36, fr->C.count, fr->C.data);
  c.vec_view(12, fr->x0.count, fr->x0.data);
  const uint8_t tag = c.get<uint8_t>();
  if (tag > 1) c.ok = false;
  fr->has_grid = tag == 1;
  if (c.ok && fr->has_grid) {
    c.vec_view(12, fr->node_ids.count, fr->node_ids.data);
    c.vec_view(4, fr->node_bits.count, fr->node_bits.data);
    c.vec_view(4, fr->node_masses.count, fr->node_masses.data);
    c.vec_view(12, fr->node_velocities.count, fr->node_velocities.data);
  }
  if (!c.ok) return fail(SVBF_FORMAT, "%s: truncated or malformed frame body", path);
  const uint64_t np = fr->flags.count;
  if (n != np || fr->energies.count != np || fr->bits.count != np || fr->x.count != np || fr->F.count != np || fr->v.count != np || fr->C.count != np || fr->x0.count != np)
    return fail(SVBF_FORMAT, "%s: particle arrays of different lengths", path);
  if (fr->has_grid) {
    const uint64_t g = fr->node_bits.count;
    if (fr->node_ids.count != g || fr->node_masses.count != g || fr->node_velocities.count != g) return fail(SVBF_FORMAT, "%s: grid arrays of different lengths", path);
  }
  *out = fr.release();
  return 0;
}
double svbf_frame_time(const SvbfFrame* f) { return f ? f->time : 0.0; }
uint64_t svbf_frame_particle_count(const SvbfFrame* f) { return f ? f->flags.count : 0; }
int64_t svbf_frame_grid_count(const SvbfFrame* f) { return f && f->has_grid ? (int64_t)f->node_bits.count : -1; }
int32_t svbf_frame_copy(const SvbfFrame* f, SvbParticles* p, SvbGrid* grid) {
  if (!f) return fail(SVBF_BAD_ARGUMENT, "svbf_frame_copy: null frame");
  auto view = [](void* dst, const SvbfFrame::View& v, size_t element_bytes) { if (dst && v.count) std::memcpy(dst, v.data, (size_t)v.count * element_bytes); };
  auto column = [](float* dst, const std::vector<float>& src) { if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * 4); };
  if (p) {
    p->n = f->flags.count;
    view(p->flags, f->flags, 4); view(p->elastic_energies, f->energies, 4); view(p->collider_bits, f->bits, 4); view(p->positions, f->x, 12);
    view(p->position_gradients, f->F, 36); view(p->velocities, f->v, 12); view(p->velocity_gradients, f->C, 36); view(p->initial_positions, f->x0, 12);
    column(p->mass, f->mass); column(p->initial_volume, f->volume); column(p->mu_or_bulk_modulus, f->p0); column(p->lambda_or_exponent, f->p1);
    column(p->sand_alpha, f->alpha); column(p->viscosity_dynamic, f->vd); column(p->viscosity_bulk, f->vb);
  }
  if (grid) {
    if (!f->has_grid) return fail(SVBF_BAD_ARGUMENT, "the frame holds no grid nodes");
    grid->n = f->node_bits.count;
    view(grid->node_ids, f->node_ids, 12); view(grid->collider_bits, f->node_bits, 4); view(grid->masses, f->node_masses, 4); view(grid->velocities, f->node_velocities, 12);
    if (grid->contributor_counts) for (uint64_t i = 0; i < grid->n; ++i) grid->contributor_counts[i] = 1;
  }
  return 0;
}
void svbf_frame_close(SvbfFrame* f) { delete f; }

// ================================================================================================ input files: reading
int32_t svbf_input_open(const char* path, const char* version, SvbfInput** out) {
  if (!path || !out) return fail(SVBF_BAD_ARGUMENT, "svbf_input_open: null argument");
  *out = nullptr;
  std::unique_ptr<SvbfInput> in(new SvbfInput());
  in->path = path;
  in->file.reset(fopen(path, "rb"));
  if (!in->file) return fail(SVBF_IO_ERROR, "failed to open %s: %s", path, std::strerror(errno));
  FILE* f = in->file.get();
  fseeko(f, 0, SEEK_END);
  in->size = (uint64_t)ftello(f);
  std::vector<uint8_t> buf;
  if (in->size < DATA_OFFSET + 8) {
    if (int rc = read_range(f, 0, in->size, buf, path)) return rc;
    Cursor c(buf.data(), buf.size());
    if (int rc = check_magic_and_version(c, INPUT_MAGIC, version, path)) return rc;
    return fail(SVBF_FORMAT, "%s: no frame index", path);
  }
  if (int rc = read_range(f, 0, DATA_OFFSET, buf, path)) return rc;
  {
    Cursor c(buf.data(), buf.size());
    if (int rc = check_magic_and_version(c, INPUT_MAGIC, version, path)) return rc;
  }
  // read_frame_offsets (reading.rs:73-81): the last 8 bytes point at the bincode Vec<u64> of frame offsets
  if (int rc = read_range(f, in->size - 8, 8, buf, path)) return rc;
  std::memcpy(&in->index_offset, buf.data(), 8);
  // an offset no seek can reach is an I/O error in the reference (InputOffsetReadingError::IoError), any other bad one ends as a bincode error
  if (in->index_offset > (uint64_t)INT64_MAX) return fail(SVBF_IO_ERROR, "%s: cannot seek to the frame index at %llu", path, (unsigned long long)in->index_offset);
  if (in->index_offset < DATA_OFFSET || in->index_offset > in->size - 8) return fail(SVBF_FORMAT, "%s: frame index offset %llu out of range", path, (unsigned long long)in->index_offset);
  if (int rc = read_range(f, in->index_offset, in->size - 8 - in->index_offset, buf, path)) return rc;
  {
    Cursor c(buf.data(), buf.size());
    if (!c.vec(in->offsets, 1)) return fail(SVBF_FORMAT, "%s: malformed frame index", path);
  }
  // read_header (reading.rs:53-58): InputHeader { consts, objects: BTreeMap<String, InputObject> } (header.rs:10-18, 73-90)
  const uint64_t header_end = in->offsets.empty() ? in->index_offset : in->offsets[0];
  if (header_end < DATA_OFFSET || header_end > in->index_offset) return fail(SVBF_FORMAT, "%s: bad first frame offset", path);
  if (int rc = read_range(f, DATA_OFFSET, header_end - DATA_OFFSET, buf, path)) return rc;
  Cursor c(buf.data(), buf.size());
  in->consts.grid_node_size = c.get<float>();
  in->consts.leaf_size = c.get<float>();
  in->consts.leaf_threshold = c.get<uint32_t>();
  in->consts.simulation_scale = c.get<float>();
  in->consts.frames_per_second = c.get<uint32_t>();
  c.take(in->consts.domain_min, 12);
  c.take(in->consts.domain_max, 12);
  const uint64_t n_objects = c.get<uint64_t>();
  for (uint64_t i = 0; c.ok && i < n_objects; ++i) {
    std::string name;
    c.string(name);
    ObjectInfo o;
    const uint32_t variant = c.get<uint32_t>();
    if (variant == 0) { o.kind = SVBF_OBJECT_PARTICLES; o.count = c.get<uint64_t>(); }
    else if (variant == 1) { o.kind = SVBF_OBJECT_COLLIDER; o.count = c.get<uint64_t>(); o.count2 = c.get<uint64_t>(); }
    else c.ok = false;
    in->objects[name] = o;
  }
  if (!c.ok) return fail(SVBF_FORMAT, "%s: truncated or malformed header", path);
  for (auto& kv : in->objects) {  // InputRanges::new (header.rs:118-160): cumulative ranges in name order
    ObjectInfo& o = kv.second;
    if (o.kind == SVBF_OBJECT_PARTICLES) { o.start = in->total_particles; in->total_particles += o.count; }
    else { o.start = in->total_vertices; o.start2 = in->total_triangles; in->total_vertices += o.count; in->total_triangles += o.count2; }
  }
  if (version) { in->version = version; in->has_version = true; }
  *out = in.release();
  return 0;
}
void svbf_input_close(SvbfInput* in) { delete in; }
uint64_t svbf_input_size(const SvbfInput* in) { return in ? in->size : 0; }
uint64_t svbf_input_frame_count(const SvbfInput* in) { return in ? in->offsets.size() : 0; }
int32_t svbf_input_consts(const SvbfInput* in, SvbConsts* out) {
  if (!in || !out) return fail(SVBF_BAD_ARGUMENT, "svbf_input_consts: null argument");
  *out = in->consts;
  return 0;
}
uint64_t svbf_input_total_particles(const SvbfInput* in) { return in ? in->total_particles : 0; }
uint64_t svbf_input_total_vertices(const SvbfInput* in) { return in ? in->total_vertices : 0; }
uint64_t svbf_input_total_triangles(const SvbfInput* in) { return in ? in->total_triangles : 0; }
uint32_t svbf_input_object_count(const SvbfInput* in) { return in ? (uint32_t)in->objects.size() : 0; }
int32_t svbf_input_object(const SvbfInput* in, uint32_t index, char* name, size_t name_cap, int32_t* kind, uint64_t* count, uint64_t* count2, uint64_t* start, uint64_t* start2) {
  if (!in || index >= in->objects.size()) return fail(SVBF_BAD_ARGUMENT, "svbf_input_object: bad index");
  auto it = in->objects.begin();
  std::advance(it, index);
  if (name && name_cap) snprintf(name, name_cap, "%s", it->first.c_str());
  if (kind) *kind = it->second.kind;
  if (count) *count = it->second.count;
  if (count2) *count2 = it->second.count2;
  if (start) *start = it->second.start;
  if (start2) *start2 = it->second.start2;
  return 0;
}

int32_t svbf_input_topology(SvbfInput* in, uint32_t* n_colliders, uint32_t* num_vertices, uint32_t* num_triangles, uint32_t* triangles) {
  if (!in) return fail(SVBF_BAD_ARGUMENT, "svbf_input_topology: null input");
  InputFrameData fr;
  if (int rc = in->read_frame(0, fr)) return rc;
  if (fr.colliders.size() > 16) return fail(SVB_TOO_MANY_COLLIDERS, "%s: %zu colliders, at most 16 are supported", in->path.c_str(), fr.colliders.size());  // frame_input.rs:168-171
  if (n_colliders) *n_colliders = (uint32_t)fr.colliders.size();
  size_t c = 0, t = 0;
  for (const auto& kv : fr.colliders) {  // name order = collider index (frame_input.rs:173-183)
    const ColliderInputData& d = kv.second;
    if (num_vertices) num_vertices[c] = (uint32_t)(d.vertex_positions.size() / 3);
    if (num_triangles) num_triangles[c] = (uint32_t)(d.triangle_indices.size() / 3);
    if (triangles && !d.triangle_indices.empty()) std::memcpy(triangles + 3 * t, d.triangle_indices.data(), d.triangle_indices.size() * 4);
    t += d.triangle_indices.size() / 3;
    ++c;
  }
  return 0;
}

// util/src/elastic.rs:33-64
static bool lame_parameters(float youngs_modulus, float poissons_ratio, float& mu, float& lambda) {
  if (youngs_modulus < 0.f) return false;
  if (!(poissons_ratio >= 0.f && poissons_ratio < 0.5f)) return false;
  mu = youngs_modulus / 2.f / (1.f + poissons_ratio);
  lambda = youngs_modulus * poissons_ratio / (1.f + poissons_ratio) / (1.f - 2.f * poissons_ratio);
  return true;
}

int32_t svbf_input_initialize(SvbfInput* in, SvbParticles* out) {
  if (!in || !out) return fail(SVBF_BAD_ARGUMENT, "svbf_input_initialize: null argument");
  InputFrameData fr;
  if (int rc = in->read_frame(0, fr)) return rc;
  const uint64_t n = in->total_particles;
  out->n = n;
  const float inv_scale = 1.f / in->consts.simulation_scale;  // initialization.rs:95-96
  // "Allocating Objects" (initialization.rs:100-128): zeros, F = identity
  auto zero = [&](auto* a, size_t k) { if (a && n) std::memset(a, 0, (size_t)n * k * sizeof(*a)); };
  zero(out->flags, 1); zero(out->mass, 1); zero(out->initial_volume, 1); zero(out->mu_or_bulk_modulus, 1); zero(out->lambda_or_exponent, 1); zero(out->sand_alpha, 1);
  zero(out->viscosity_dynamic, 1); zero(out->viscosity_bulk, 1); zero(out->initial_positions, 3); zero(out->positions, 3); zero(out->position_gradients, 9);
  zero(out->velocities, 3); zero(out->velocity_gradients, 9); zero(out->elastic_energies, 1); zero(out->collider_bits, 1);
  if (out->position_gradients)
    for (uint64_t i = 0; i < n; ++i) out->position_gradients[9 * i] = out->position_gradients[9 * i + 4] = out->position_gradients[9 * i + 8] = 1.f;
  for (const auto& kv : fr.particles) {
    const std::string& name = kv.first;
    const ParticlesInputData& d = kv.second;
    auto it = in->objects.find(name);
    if (it == in->objects.end()) return fail(SVBF_OBJECT_ERROR, "The object is missing in the header: %s", name.c_str());
    if (it->second.kind != SVBF_OBJECT_PARTICLES) return fail(SVBF_OBJECT_ERROR, "The object's type doesn't match the one in the header: %s", name.c_str());
    const uint64_t first = it->second.start, count = it->second.count;
    if (!d.has_transforms) return fail(SVBF_MISSING_INPUT, "'%s': missing input for input_transforms", name.c_str());
    if (!d.has_sizes) return fail(SVBF_MISSING_INPUT, "'%s': missing input for input_sizes", name.c_str());
    if (!d.has_densities) return fail(SVBF_MISSING_INPUT, "'%s': missing input for input_densities", name.c_str());
    // the reference slices by the header's range and indexes the inputs by particle: shorter inputs are an error there (panic)
    if (d.flags.size() != count || d.transforms.size() != 16 * count || d.sizes.size() != count || d.densities.size() != count)
      return fail(SVBF_LENGTH_MISMATCH, "'%s': expected %llu values per attribute", name.c_str(), (unsigned long long)count);
    auto short_of = [&](bool has, size_t len, size_t per) { return has && len != per * count; };
    if (short_of(d.has_youngs, d.youngs.size(), 1) || short_of(d.has_poissons, d.poissons.size(), 1) || short_of(d.has_visc_dynamic, d.visc_dynamic.size(), 1) ||
        short_of(d.has_visc_bulk, d.visc_bulk.size(), 1) || short_of(d.has_exponents, d.exponents.size(), 1) || short_of(d.has_bulk, d.bulk.size(), 1) ||
        short_of(d.has_sand, d.sand.size(), 1) || short_of(d.has_initial_positions, d.initial_positions.size(), 3) || short_of(d.has_initial_velocities, d.initial_velocities.size(), 3))
      return fail(SVBF_LENGTH_MISMATCH, "'%s': expected %llu values per attribute", name.c_str(), (unsigned long long)count);
    for (uint64_t k = 0; k < count; ++k) {
      const uint64_t i = first + k;
      const uint32_t fl = d.flags[k];
      auto invalid = [&](const char* what) { return fail(SVBF_PARTICLE_INVALID, "'%s': input particle #%llu invalid: %s", name.c_str(), (unsigned long long)k, what); };
      auto missing = [&](const char* attribute) {
        return fail(SVBF_MISSING_INPUT, "'%s': input particle #%llu invalid: This setup requires some input for %s", name.c_str(), (unsigned long long)k, attribute);
      };
      if (fl & ~F_ALL) return invalid("Some flags are set that are not know");
      if (((fl & F_IS_SOLID) != 0) == ((fl & F_IS_FLUID) != 0)) return invalid("The particle solid or fluid flag must be set, but not both");
      if (out->flags) out->flags[i] = fl;
      const float s = inv_scale * d.sizes[k];
      const float volume = s * s * s;                                    // (inv_scale * size).powi(3)
      if (out->initial_volume) out->initial_volume[i] = volume;
      if (out->mass) out->mass[i] = volume * d.densities[k];
      if (fl & F_USE_VISCOSITY) {
        if (!d.has_visc_dynamic) return missing("input_viscosities_dynamic");
        if (!d.has_visc_bulk) return missing("input_viscosities_bulk");
        if (out->viscosity_dynamic) out->viscosity_dynamic[i] = d.visc_dynamic[k];
        if (out->viscosity_bulk) out->viscosity_bulk[i] = d.visc_bulk[k];
      }
      if (fl & F_IS_SOLID) {
        if (!d.has_youngs) return missing("input_youngs_moduluses");
        if (!d.has_poissons) return missing("input_poissons_ratios");
        float mu, lambda;
        if (!lame_parameters(d.youngs[k], d.poissons[k], mu, lambda)) return invalid("Energy error: Young's modulus or Poisson's ratio out of bounds");
        if (out->mu_or_bulk_modulus) out->mu_or_bulk_modulus[i] = mu;
        if (out->lambda_or_exponent) out->lambda_or_exponent[i] = lambda;
        if (fl & F_USE_SAND_ALPHA) {
          if (!d.has_sand) return missing("input_sand_alphas");
          if (out->sand_alpha) out->sand_alpha[i] = d.sand[k];
        }
      } else {
        if (!d.has_exponents) return missing("input_exponents");
        if (!d.has_bulk) return missing("input_bulk_moduluses");
        const int32_t exponent = (int32_t)d.exponents[k];
        if (!(exponent > 1)) return invalid("Energy error: exponent out of bounds");           // elastic.rs:533-540
        if (d.bulk[k] < 0.f) return invalid("Energy error: bulk modulus out of bounds");       // elastic.rs:524-531
        if (out->mu_or_bulk_modulus) out->mu_or_bulk_modulus[i] = d.bulk[k];
        if (out->lambda_or_exponent) out->lambda_or_exponent[i] = (float)exponent;
      }
      const float* t = &d.transforms[16 * k];  // [[f32;4];4]: rows 0..2 = the columns of F, row 3 = translation (initialization.rs:243-260)
      if (out->position_gradients)
        for (int c = 0; c < 3; ++c)
          for (int r = 0; r < 3; ++r) out->position_gradients[9 * i + 3 * c + r] = t[4 * c + r];
      if (out->positions)
        for (int r = 0; r < 3; ++r) out->positions[3 * i + r] = inv_scale * t[12 + r];
    }
    if (d.has_initial_velocities && out->velocities && count) std::memcpy(out->velocities + 3 * first, d.initial_velocities.data(), (size_t)count * 12);
    if (d.has_initial_positions && out->initial_positions && count) std::memcpy(out->initial_positions + 3 * first, d.initial_positions.data(), (size_t)count * 12);
  }
  return 0;
}

int32_t svbf_input_keyframe(SvbfInput* in, uint64_t frame, float gravity[3], uint32_t* particle_flags, float* goals, float* vertex_positions, float* frictions, float* dampings) {
  if (!in) return fail(SVBF_BAD_ARGUMENT, "svbf_input_keyframe: null input");
  InputFrameData fr;
  if (int rc = in->read_frame(frame, fr)) return rc;
  const uint64_t n = in->total_particles;
  const float scale = in->consts.simulation_scale;
  if (gravity) std::memcpy(gravity, fr.gravity, 12);
  if (particle_flags && n) std::memset(particle_flags, 0, (size_t)n * 4);
  if (goals && n) std::memset(goals, 0, (size_t)n * 12);
  for (const auto& kv : fr.particles) {  // frame_input.rs:82-108
    auto it = in->objects.find(kv.first);
    if (it == in->objects.end()) return fail(SVBF_OBJECT_ERROR, "object not in header: %s", kv.first.c_str());
    if (it->second.kind != SVBF_OBJECT_PARTICLES) return fail(SVBF_OBJECT_ERROR, "object changed type: %s", kv.first.c_str());
    const ParticlesInputData& d = kv.second;
    if (d.flags.size() != it->second.count) return fail(SVBF_LENGTH_MISMATCH, "'%s': %zu flags for %llu particles", kv.first.c_str(), d.flags.size(), (unsigned long long)it->second.count);
    if (particle_flags && !d.flags.empty()) std::memcpy(particle_flags + it->second.start, d.flags.data(), d.flags.size() * 4);
    if (d.has_goals) {
      if (d.goals.size() != 3 * d.flags.size())
        return fail(SVBF_LENGTH_MISMATCH, "'%s': length mismatch between 'Particle Flags' and 'Particle Goal Positions'", kv.first.c_str());
      if (goals)
        for (size_t q = 0; q < d.goals.size(); ++q) goals[3 * it->second.start + q] = d.goals[q] / scale;  // frame_input.rs:118-120
    }
  }
  size_t v = 0, t = 0;
  for (const auto& kv : fr.colliders) {  // into_values(): name order (frame_input.rs:110-117)
    const ColliderInputData& d = kv.second;
    if (vertex_positions)
      for (size_t q = 0; q < d.vertex_positions.size(); ++q) vertex_positions[3 * v + q] = d.vertex_positions[q] / scale;  // :121-123
    if (frictions && !d.frictions.empty()) std::memcpy(frictions + t, d.frictions.data(), d.frictions.size() * 4);
    if (dampings && !d.dampings.empty()) std::memcpy(dampings + t, d.dampings.data(), d.dampings.size() * 4);
    v += d.vertex_positions.size() / 3;
    t += d.frictions.size();
  }
  return 0;
}

// ================================================================================================ input files: recording
int32_t svbf_input_writer_open(const char* path, const char* version, const SvbConsts* consts, const SvbfObjectDesc* objects, uint32_t n_objects, SvbfInputWriter** out) {
  if (!path || !consts || !out || (n_objects && !objects)) return fail(SVBF_BAD_ARGUMENT, "svbf_input_writer_open: null argument");
  *out = nullptr;
  uint8_t ver[VERSION_LEN];
  if (!version_bytes(version, ver)) return fail(SVBF_BAD_ARGUMENT, "version string too long");
  std::unique_ptr<SvbfInputWriter> w(new SvbfInputWriter());
  w->path = path;
  for (uint32_t i = 0; i < n_objects; ++i) {
    if (!objects[i].name) return fail(SVBF_BAD_ARGUMENT, "object %u has no name", i);
    ObjectInfo o;
    o.kind = objects[i].kind;
    o.count = objects[i].count;
    o.count2 = objects[i].count2;
    w->objects[objects[i].name] = o;
  }
  w->file.reset(fopen(path, "wb"));
  if (!w->file) return fail(SVBF_IO_ERROR, "failed to create %s: %s", path, std::strerror(errno));
  Sink& s = w->sink;
  s.f = w->file.get();
  s.put(INPUT_MAGIC, MAGIC_LEN);
  s.put(ver, VERSION_LEN);
  s.val<float>(consts->grid_node_size);
  s.val<float>(consts->leaf_size);
  s.val<uint32_t>(consts->leaf_threshold);
  s.val<float>(consts->simulation_scale);
  s.val<uint32_t>(consts->frames_per_second);
  s.put(consts->domain_min, 12);
  s.put(consts->domain_max, 12);
  s.val<uint64_t>(w->objects.size());
  for (const auto& kv : w->objects) {
    s.string(kv.first);
    if (kv.second.kind == SVBF_OBJECT_PARTICLES) { s.val<uint32_t>(0); s.val<uint64_t>(kv.second.count); }
    else { s.val<uint32_t>(1); s.val<uint64_t>(kv.second.count); s.val<uint64_t>(kv.second.count2); }
  }
  if (!s.ok) return fail(SVBF_IO_ERROR, "failed to write the header of %s", path);
  *out = w.release();
  return 0;
}

int32_t svbf_input_writer_frame(SvbfInputWriter* w, const float gravity[3], const SvbfParticlesInput* particles, uint32_t n_particles_inputs, const SvbfColliderInput* colliders,
                                uint32_t n_collider_inputs) {
  if (!w || !gravity || (n_particles_inputs && !particles) || (n_collider_inputs && !colliders)) return fail(SVBF_BAD_ARGUMENT, "svbf_input_writer_frame: null argument");
  if (n_collider_inputs > 16) return fail(SVB_TOO_MANY_COLLIDERS, "too many colliders");  // collider_inputs.rs:48-61
  std::map<std::string, const SvbfParticlesInput*> pm;
  std::map<std::string, const SvbfColliderInput*> cm;
  for (uint32_t i = 0; i < n_particles_inputs; ++i) pm[particles[i].name ? particles[i].name : ""] = &particles[i];
  for (uint32_t i = 0; i < n_collider_inputs; ++i) cm[colliders[i].name ? colliders[i].name : ""] = &colliders[i];
  // InputFrame::verify (frame.rs:66-187)
  for (const auto& kv : pm) {
    auto it = w->objects.find(kv.first);
    if (it == w->objects.end()) return fail(SVBF_OBJECT_ERROR, "frame %zu: object not in header: %s", w->offsets.size(), kv.first.c_str());
    if (it->second.kind != SVBF_OBJECT_PARTICLES) return fail(SVBF_OBJECT_ERROR, "frame %zu: object changed type: %s", w->offsets.size(), kv.first.c_str());
    if (kv.second->n != it->second.count)
      return fail(SVBF_LENGTH_MISMATCH, "frame %zu: '%s': found %llu, expected %llu", w->offsets.size(), kv.first.c_str(), (unsigned long long)kv.second->n, (unsigned long long)it->second.count);
    if (kv.second->n && !kv.second->flags) return fail(SVBF_BAD_ARGUMENT, "'%s': flags are required", kv.first.c_str());
  }
  for (const auto& kv : cm) {
    auto it = w->objects.find(kv.first);
    if (it == w->objects.end()) return fail(SVBF_OBJECT_ERROR, "frame %zu: object not in header: %s", w->offsets.size(), kv.first.c_str());
    if (it->second.kind != SVBF_OBJECT_COLLIDER) return fail(SVBF_OBJECT_ERROR, "frame %zu: object changed type: %s", w->offsets.size(), kv.first.c_str());
  }
  for (const auto& kv : w->objects)
    if (kv.second.kind == SVBF_OBJECT_COLLIDER) {
      auto it = cm.find(kv.first);
      if (it == cm.end()) return fail(SVBF_COLLIDER_INPUT_MISSING, "frame %zu: collider input missing: %s", w->offsets.size(), kv.first.c_str());
      if (it->second->num_vertices != kv.second.count || it->second->num_triangles != kv.second.count2)
        return fail(SVBF_LENGTH_MISMATCH, "frame %zu: '%s': collider sizes differ from the header", w->offsets.size(), kv.first.c_str());
    }
  Sink& s = w->sink;
  w->offsets.push_back(s.pos);  // record_frame (writing.rs:40-50)
  s.put(gravity, 12);
  s.val<uint64_t>(pm.size());
  for (const auto& kv : pm) {
    const SvbfParticlesInput& d = *kv.second;
    s.string(kv.first);
    s.vec(d.flags, d.n, 1);
    s.opt_vec(d.transforms, d.n, 16);
    s.opt_vec(d.sizes, d.n, 1);
    s.opt_vec(d.densities, d.n, 1);
    s.opt_vec(d.youngs_moduluses, d.n, 1);
    s.opt_vec(d.poissons_ratios, d.n, 1);
    s.opt_vec(d.initial_positions, d.n, 3);
    s.opt_vec(d.initial_velocities, d.n, 3);
    s.opt_vec(d.viscosities_dynamic, d.n, 1);
    s.opt_vec(d.viscosities_bulk, d.n, 1);
    s.opt_vec(d.exponents, d.n, 1);
    s.opt_vec(d.bulk_moduluses, d.n, 1);
    s.opt_vec(d.sand_alphas, d.n, 1);
    s.opt_vec(d.goal_positions, d.n, 3);
  }
  s.val<uint64_t>(cm.size());
  for (const auto& kv : cm) {
    const SvbfColliderInput& d = *kv.second;
    s.string(kv.first);
    s.vec(d.vertex_positions, d.num_vertices, 3);
    s.vec(d.triangle_indices, d.num_triangles, 3);
    s.vec(d.triangle_frictions, d.num_triangles, 1);
    s.vec(d.triangle_dampings, d.num_triangles, 1);
  }
  return s.ok ? 0 : fail(SVBF_IO_ERROR, "failed to write frame %zu of %s", w->offsets.size() - 1, w->path.c_str());
}

int32_t svbf_input_writer_finish(SvbfInputWriter* w) {
  if (!w) return fail(SVBF_BAD_ARGUMENT, "svbf_input_writer_finish: null writer");
  std::unique_ptr<SvbfInputWriter> guard(w);
  Sink& s = w->sink;
  const uint64_t index_offset = s.pos;  // flush (writing.rs:52-68)
  s.vec(w->offsets.data(), w->offsets.size(), 1);
  s.val<uint64_t>(index_offset);
  const bool flushed = fflush(s.f) == 0;
  return s.ok && flushed ? 0 : fail(SVBF_IO_ERROR, "failed to finish %s", w->path.c_str());
}

}  // extern "C"
