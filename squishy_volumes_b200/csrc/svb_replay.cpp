// svb_replay — the compute thread's frame loop (rust/crates/core/src/compute_thread.rs:60-190) over the two C ABIs:
// reads <cache_dir>/simulation_input.bin, writes <cache_dir>/frame_%05d.bin, one frame file per output frame, exactly the
// files the reference's cache (rust/crates/cache/src/cache.rs) serves to Blender.  It stands in for the ~100-line Rust
// `ComputeState::B200` arm of INTEGRATION.md, which cannot be compiled in this image (no cargo).
//
//   svb_replay <cache_dir> <number_of_frames> <max_time_step> [--adaptive] [--device N] [--next-frame K] [--no-grid]
//
// Exit status: 0 done, 2 simulation-level error (the failing frame is stored first, compute_thread.rs:165-169), 1 fatal.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/svb200.h"
#include "../../include/svb_files.h"

namespace {

struct ParticleArrays {  // owns the arrays an SvbParticles points at
  std::vector<uint32_t> flags, bits;
  std::vector<float> mass, volume, p0, p1, alpha, vd, vb, x0, x, F, v, C, energy;
  SvbParticles view(uint64_t n) {
    flags.resize(n); bits.resize(n); mass.resize(n); volume.resize(n); p0.resize(n); p1.resize(n); alpha.resize(n); vd.resize(n); vb.resize(n);
    x0.resize(3 * n); x.resize(3 * n); F.resize(9 * n); v.resize(3 * n); C.resize(9 * n); energy.resize(n);
    SvbParticles s{};
    s.n = n;
    s.flags = flags.data(); s.mass = mass.data(); s.initial_volume = volume.data(); s.mu_or_bulk_modulus = p0.data(); s.lambda_or_exponent = p1.data();
    s.sand_alpha = alpha.data(); s.viscosity_dynamic = vd.data(); s.viscosity_bulk = vb.data(); s.initial_positions = x0.data(); s.positions = x.data();
    s.position_gradients = F.data(); s.velocities = v.data(); s.velocity_gradients = C.data(); s.elastic_energies = energy.data(); s.collider_bits = bits.data();
    return s;
  }
};
struct GridArrays {
  std::vector<int32_t> ids;
  std::vector<uint32_t> bits;
  std::vector<float> masses, velocities;
  SvbGrid view(uint64_t n) {
    ids.resize(3 * n); bits.resize(n); masses.resize(n); velocities.resize(3 * n);
    SvbGrid g{};
    g.n = n;
    g.node_ids = ids.data(); g.collider_bits = bits.data(); g.masses = masses.data(); g.velocities = velocities.data(); g.contributor_counts = nullptr;
    return g;
  }
};
struct KeyframeArrays {
  std::vector<uint32_t> flags;
  std::vector<float> goals, vertices, frictions, dampings;
  SvbKeyframe k{};
  int load(SvbfInput* in, uint64_t frame) {
    const uint64_t n = svbf_input_total_particles(in), nv = svbf_input_total_vertices(in), nt = svbf_input_total_triangles(in);
    flags.assign(n, 0); goals.assign(3 * n, 0.f); vertices.assign(3 * nv, 0.f); frictions.assign(nt, 0.f); dampings.assign(nt, 0.f);
    if (int rc = svbf_input_keyframe(in, frame, k.gravity, flags.data(), goals.data(), vertices.data(), frictions.data(), dampings.data())) return rc;
    k.particle_flags = flags.data(); k.particle_goal_positions = goals.data(); k.vertex_positions = vertices.data();
    k.triangle_frictions = frictions.data(); k.triangle_dampings = dampings.data();
    return 0;
  }
};

// The reference hands finished frames to a store thread (rust/crates/cache/src/store_thread.rs) so that serialising and
// writing ~170 bytes per particle does not stall the simulation; same here: the frame loop downloads into a job, the worker
// writes it (svbf_frame_write: temp.bin + rename, one writer at a time), at most two frames wait in the queue.
struct FrameJob {
  uint64_t frame = 0;
  double time = 0;
  uint64_t n = 0, ng = 0;
  bool has_grid = false;
  ParticleArrays pa;
  GridArrays ga;
};
class StoreThread {
 public:
  explicit StoreThread(std::string dir) : dir_(std::move(dir)), worker_([this] { run(); }) {}
  ~StoreThread() { finish(); }
  // blocks while two frames are already waiting (back-pressure instead of unbounded host memory)
  void push(std::unique_ptr<FrameJob> job) {
    std::unique_lock<std::mutex> lock(m_);
    room_.wait(lock, [this] { return queue_.size() < 2; });
    queue_.push_back(std::move(job));
    work_.notify_one();
  }
  // waits for everything queued; returns false (and the message) when a frame could not be stored
  bool finish(std::string* message = nullptr) {
    {
      std::unique_lock<std::mutex> lock(m_);
      done_ = true;
      work_.notify_one();
    }
    if (worker_.joinable()) worker_.join();
    if (message) *message = error_;
    return error_.empty();
  }

 private:
  void run() {
    for (;;) {
      std::unique_ptr<FrameJob> job;
      {
        std::unique_lock<std::mutex> lock(m_);
        work_.wait(lock, [this] { return done_ || !queue_.empty(); });
        if (queue_.empty()) return;
        job = std::move(queue_.front());
        queue_.pop_front();
        room_.notify_one();
      }
      char path[4096];
      svbf_frame_path(dir_.c_str(), job->frame, path, sizeof path);
      SvbParticles p = job->pa.view(job->n);
      SvbGrid g = job->ga.view(job->ng);
      uint64_t bytes = 0;
      if (svbf_frame_write(path, nullptr, job->time, &p, job->has_grid ? &g : nullptr, &bytes)) {
        std::lock_guard<std::mutex> lock(m_);
        if (error_.empty()) error_ = svbf_last_error();
      } else std::printf("stored frame %llu: %llu bytes\n", (unsigned long long)job->frame, (unsigned long long)bytes);
    }
  }
  std::string dir_;
  std::mutex m_;
  std::condition_variable work_, room_;
  std::deque<std::unique_ptr<FrameJob>> queue_;
  bool done_ = false;
  std::string error_;
  std::thread worker_;
};

int die(const char* what, const char* detail) {
  std::fprintf(stderr, "svb_replay: %s: %s\n", what, detail ? detail : "");
  return 1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: svb_replay <cache_dir> <number_of_frames> <max_time_step> [--adaptive] [--device N] [--next-frame K] [--no-grid]\n");
    return 1;
  }
  const std::string cache_dir = argv[1];
  const uint64_t number_of_frames = std::strtoull(argv[2], nullptr, 10);
  const float max_time_step = std::strtof(argv[3], nullptr);
  int adaptive = 0, device = 0, store_grid = 1;
  uint64_t next_frame = 0;
  for (int i = 4; i < argc; ++i) {
    if (!std::strcmp(argv[i], "--adaptive")) adaptive = 1;
    else if (!std::strcmp(argv[i], "--no-grid")) store_grid = 0;
    else if (!std::strcmp(argv[i], "--device") && i + 1 < argc) device = std::atoi(argv[++i]);
    else if (!std::strcmp(argv[i], "--next-frame") && i + 1 < argc) next_frame = std::strtoull(argv[++i], nullptr, 10);
    else return die("unknown argument", argv[i]);
  }
  // compute_thread.rs:64-69: the recorded input next to the frames
  SvbfInput* in = nullptr;
  if (svbf_input_open((cache_dir + "/simulation_input.bin").c_str(), nullptr, &in)) return die("input", svbf_last_error());
  SvbConsts consts{};
  svbf_input_consts(in, &consts);
  const uint64_t n = svbf_input_total_particles(in), recorded = svbf_input_frame_count(in);
  if (recorded == 0) return die("input", "the input file holds no frames");
  char path[4096];
  ParticleArrays pa;
  SvbParticles particles = pa.view(n);
  double time = 0.0;
  if (next_frame == 0) {  // compute_thread.rs:81-89: create the initial state and store it as frame 0
    if (svbf_input_initialize(in, &particles)) return die("initialize_io_state", svbf_last_error());
    GridArrays empty;
    SvbGrid g = empty.view(0);
    svbf_frame_path(cache_dir.c_str(), 0, path, sizeof path);
    if (svbf_frame_write(path, nullptr, 0.0, &particles, &g, nullptr)) return die("store frame 0", svbf_last_error());
    next_frame = 1;
  } else {  // :90-96: resume from the last stored frame
    SvbfFrame* f = nullptr;
    svbf_frame_path(cache_dir.c_str(), next_frame - 1, path, sizeof path);
    if (svbf_frame_open(path, nullptr, &f)) return die("checkpoint", svbf_last_error());
    if (svbf_frame_particle_count(f) != n) return die("checkpoint", "particle count differs from the input header");
    time = svbf_frame_time(f);
    svbf_frame_copy(f, &particles, nullptr);
    svbf_frame_close(f);
  }
  // :98-117: FrameInput::new + from_io_state.  There is no CPU arm: without a CUDA device this fails.
  uint32_t n_colliders = 0;
  if (svbf_input_topology(in, &n_colliders, nullptr, nullptr, nullptr)) return die("topology", svbf_last_error());
  std::vector<uint32_t> nv(n_colliders + 1), nt(n_colliders + 1), tris(3 * svbf_input_total_triangles(in) + 3);
  if (svbf_input_topology(in, &n_colliders, nv.data(), nt.data(), tris.data())) return die("topology", svbf_last_error());
  SvbHandle* h = nullptr;
  if (int rc = svb_create(&consts, &particles, time, device, &h)) {
    std::fprintf(stderr, "svb_replay: svb_create failed (%d): %s\n", rc, h ? svb_last_error(h) : "no CUDA device (there is no CPU fallback)");
    if (h) svb_destroy(h);
    return 1;
  }
  if (svb_set_topology(h, n_colliders, nv.data(), nt.data(), tris.data())) return die("svb_set_topology", svb_last_error(h));
  svb_set_option(h, "store_grid", store_grid ? 1.0 : 0.0);
  KeyframeArrays ka, kb;
  uint64_t loaded_a = ~0ull;
  int status = 0;
  StoreThread store(cache_dir);
  while (next_frame < number_of_frames) {  // :125-177
    const auto t0 = std::chrono::steady_clock::now();
    const uint64_t frame = next_frame - 1;
    const uint64_t a = frame < recorded - 1 ? frame : recorded - 1;  // load_points, frame_input.rs:289-293: past the end the last keyframe stays
    if (a != loaded_a) {
      if (ka.load(in, a)) return die("keyframe", svbf_last_error());
      if (a + 1 < recorded && kb.load(in, a + 1)) return die("keyframe", svbf_last_error());
      loaded_a = a;
    }
    if (svb_set_keyframes(h, frame, &ka.k, a + 1 < recorded ? &kb.k : nullptr)) return die("svb_set_keyframes", svb_last_error(h));
    const double target_time = (double)next_frame / (double)consts.frames_per_second;
    const uint64_t substeps_before = svb_substeps(h);
    const int rc = svb_advance(h, target_time, max_time_step, adaptive, nullptr, nullptr, nullptr);
    if (rc < 0) return die("svb_advance", svb_last_error(h));
    std::unique_ptr<FrameJob> job(new FrameJob());
    job->frame = next_frame;
    job->time = svb_time(h);
    job->n = n;
    SvbParticles out = job->pa.view(n);
    if (svb_download(h, &out)) return die("svb_download", svb_last_error(h));
    if (store_grid) {
      const int64_t count = svb_grid_count(h);
      if (count < 0) return die("svb_grid_count", svb_last_error(h));
      job->has_grid = true;
      job->ng = (uint64_t)count;
      SvbGrid g = job->ga.view(job->ng);
      if (svb_download_grid(h, &g)) return die("svb_download_grid", svb_last_error(h));
    }
    store.push(std::move(job));  // stored even if the substep loop failed (compute_thread.rs:165-169)
    const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("frame %llu of %llu: %llu substeps, %.3f s%s\n", (unsigned long long)next_frame, (unsigned long long)number_of_frames,
                (unsigned long long)(svb_substeps(h) - substeps_before), seconds, rc > 0 ? " (simulation error)" : "");
    if (rc > 0) {
      std::fprintf(stderr, "svb_replay: simulation-level error %d at frame %llu: %s\n", rc, (unsigned long long)next_frame, svb_last_error(h));
      status = 2;
      break;
    }
    ++next_frame;
  }
  std::string store_error;
  if (!store.finish(&store_error)) return die("store frame", store_error.c_str());
  svb_destroy(h);
  svbf_input_close(in);
  return status;
}
