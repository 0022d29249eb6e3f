// svb_multi.inl — one handle over several GPUs of the box, driven from ONE caller thread (included at the end of svb200.cu).
//
// The reference drives a back end from a single compute thread (core/src/compute_thread.rs:78,100-163), so a drop-in that spans
// the 8 GPUs of a box cannot ask for one process per GPU.  `svb_create_multi` returns an ordinary SvbHandle whose calls fan out to
// one slab rank per device, each driven by a short-lived worker thread of the library for the duration of the call:
//   * the particle state is cut into slabs along x on grid-block planes by particle count (what squishy_volumes_b200/slabs.py does
//     for the process-per-GPU launcher), split on upload and re-assembled in original particle order on download;
//   * the ranks' mailboxes are plain device allocations reached through cudaDeviceEnablePeerAccess — same process, so no CUDA IPC
//     and no NCCL bootstrap; the exchange kernels (halo sums, migrating rows, adaptive-step limits) are the ones of the
//     process-per-GPU path, reading and writing the neighbour's HBM over NVLink.
#include <thread>

struct SvbMulti {
  std::vector<SvbHandle*> ranks;
  std::vector<int> devices;
  std::vector<std::pair<int, int>> plan;   // block columns [lo, hi) per rank
  uint64_t n_global = 0;
  SvbConsts consts{};
};

namespace {

constexpr int MULTI_FAR = 1 << 15;   // "infinity" in block units, inside the +-2^16 key range (slabs.py: FAR)

// block column of a position: floor(floor(x/h - 1/2) / 4) with the device's f32 sequence (svb_kernels.cuh: base_node)
inline int host_block_x(float x, float h) {
  const float q = x / h;
  const float t = q - 0.5f;
  const int cell = (int)std::floor(t);
  return cell >> 2;
}

// cuts on block planes that balance the particle count; the outer slabs extend to +-FAR (slabs.py: plan_slabs)
std::vector<std::pair<int, int>> plan_slabs_host(const float* positions, uint64_t n, float h, int n_ranks) {
  std::vector<int> cuts;
  if (n == 0) {
    for (int r = 1; r < n_ranks; ++r) cuts.push_back(r);
  } else {
    int lo = INT32_MAX, hi = INT32_MIN;
    std::vector<int> bx(n);
    for (uint64_t i = 0; i < n; ++i) {
      bx[i] = host_block_x(positions[3 * i], h);
      lo = std::min(lo, bx[i]);
      hi = std::max(hi, bx[i]);
    }
    std::vector<uint64_t> hist((size_t)(hi - lo + 1), 0);
    for (uint64_t i = 0; i < n; ++i) ++hist[(size_t)(bx[i] - lo)];
    std::vector<uint64_t> cum(hist.size());
    uint64_t run = 0;
    for (size_t k = 0; k < hist.size(); ++k) { run += hist[k]; cum[k] = run; }
    int prev = lo;
    for (int r = 1; r < n_ranks; ++r) {
      const double target = (double)n * r / n_ranks;
      size_t k = 0;
      while (k + 1 < cum.size() && (double)cum[k] < target) ++k;   // first column whose cumulative count reaches the target
      int cut = lo + (int)k + 1;                                     // cut after column k
      cut = std::max(cut, prev + 1);
      cuts.push_back(cut);
      prev = cut;
    }
  }
  std::vector<std::pair<int, int>> plan;
  int from = -MULTI_FAR;
  for (int r = 0; r < n_ranks; ++r) {
    const int to = r + 1 < n_ranks ? cuts[(size_t)r] : MULTI_FAR;
    plan.emplace_back(from, to);
    from = to;
  }
  return plan;
}

struct HostRows {   // one rank's share of an SvbParticles, gathered into contiguous host arrays
  std::vector<uint32_t> index;   // original (global) index of every row
  std::vector<uint32_t> flags, bits;
  std::vector<float> mass, vol, p0, p1, alpha, vd, vb, x0, x, F, v, C, energy;
  SvbParticles view() {
    SvbParticles p{};
    p.n = index.size();
    p.flags = flags.data(); p.mass = mass.data(); p.initial_volume = vol.data(); p.mu_or_bulk_modulus = p0.data(); p.lambda_or_exponent = p1.data();
    p.sand_alpha = alpha.data(); p.viscosity_dynamic = vd.data(); p.viscosity_bulk = vb.data(); p.initial_positions = x0.data(); p.positions = x.data();
    p.position_gradients = F.data(); p.velocities = v.data(); p.velocity_gradients = C.data(); p.elastic_energies = energy.data(); p.collider_bits = bits.data();
    return p;
  }
  void resize(size_t m) {
    index.resize(m); flags.resize(m); bits.resize(m); mass.resize(m); vol.resize(m); p0.resize(m); p1.resize(m); alpha.resize(m); vd.resize(m); vb.resize(m);
    x0.resize(3 * m); x.resize(3 * m); F.resize(9 * m); v.resize(3 * m); C.resize(9 * m); energy.resize(m);
  }
};

void gather_rows(const SvbParticles* p, const std::vector<uint32_t>& idx, HostRows& out) {
  const size_t m = idx.size();
  out.resize(m);
  out.index = idx;
  auto take1 = [&](const float* src, std::vector<float>& dst) { for (size_t q = 0; q < m; ++q) dst[q] = src ? src[idx[q]] : 0.f; };
  auto takek = [&](const float* src, std::vector<float>& dst, int k) {
    for (size_t q = 0; q < m; ++q)
      for (int c = 0; c < k; ++c) dst[q * k + c] = src ? src[(size_t)idx[q] * k + c] : 0.f;
  };
  for (size_t q = 0; q < m; ++q) { out.flags[q] = p->flags ? p->flags[idx[q]] : 0u; out.bits[q] = p->collider_bits ? p->collider_bits[idx[q]] : 0u; }
  take1(p->mass, out.mass); take1(p->initial_volume, out.vol); take1(p->mu_or_bulk_modulus, out.p0); take1(p->lambda_or_exponent, out.p1);
  take1(p->sand_alpha, out.alpha); take1(p->viscosity_dynamic, out.vd); take1(p->viscosity_bulk, out.vb); take1(p->elastic_energies, out.energy);
  takek(p->initial_positions, out.x0, 3); takek(p->positions, out.x, 3); takek(p->velocities, out.v, 3);
  takek(p->position_gradients, out.F, 9); takek(p->velocity_gradients, out.C, 9);
}

template <class Fn>
int for_each_rank(SvbMulti* m, Fn fn) {   // fn(rank) -> status; one worker thread per device for the duration of the call
  const int n = (int)m->ranks.size();
  std::vector<int> rc((size_t)n, 0);
  std::vector<std::thread> workers;
  for (int r = 1; r < n; ++r) workers.emplace_back([&, r] { rc[(size_t)r] = fn(r); });
  rc[0] = fn(0);
  for (auto& w : workers) w.join();
  int worst = 0;
  for (int r = 0; r < n; ++r) {
    if (rc[(size_t)r] < 0 && (worst >= 0 || rc[(size_t)r] < worst)) worst = rc[(size_t)r];
    else if (rc[(size_t)r] > 0 && worst >= 0) worst |= rc[(size_t)r];
  }
  return worst;
}

void multi_note_error(SvbHandle* front, int rc) {
  if (rc == 0) return;
  for (SvbHandle* r : front->multi->ranks)
    if (!r->last_error.empty()) { front->last_error = "device " + std::to_string(r->device) + ": " + r->last_error; break; }
}

// split `p` by the plan and (re)load every rank; `create`: the rank handles do not exist yet
int multi_load(SvbHandle* front, const SvbParticles* p, double time, bool create) {
  SvbMulti* m = front->multi;
  const int n_ranks = (int)m->devices.size();
  const float h = m->consts.grid_node_size / m->consts.simulation_scale;
  if (p->n && !p->positions) return fail(front, SVB_BAD_ARGUMENT, "a required particle array is NULL");
  m->n_global = p->n;
  m->plan = plan_slabs_host(p->positions, p->n, h, n_ranks);
  std::vector<std::vector<uint32_t>> owner_rows((size_t)n_ranks);
  for (uint64_t i = 0; i < p->n; ++i) {
    const int bx = host_block_x(p->positions[3 * i], h);
    int r = 0;
    while (r + 1 < n_ranks && bx >= m->plan[(size_t)r].second) ++r;
    owner_rows[(size_t)r].push_back((uint32_t)i);
  }
  size_t n_max = 0;
  for (auto& rows : owner_rows) n_max = std::max(n_max, rows.size());
  if (create) m->ranks.assign((size_t)n_ranks, nullptr);
  int rc = for_each_rank(m, [&](int r) -> int {
    HostRows rows;
    gather_rows(p, owner_rows[(size_t)r], rows);
    SvbParticles view = rows.view();
    SvbHandle*& h_r = m->ranks[(size_t)r];
    int rc_r = create ? svb_create(&m->consts, &view, time, m->devices[(size_t)r], &h_r) : svb_upload(h_r, &view, time);
    if (rc_r) return rc_r;
    h_r->n_global = (uint32_t)p->n;
    if (!rows.index.empty()) rc_r = svb_set_original_indices(h_r, rows.index.data(), rows.index.size());
    return rc_r;
  });
  if (rc) { multi_note_error(front, rc); return rc; }
  if (!create) {   // the ranks keep their mailboxes; only the slab table changes
    for (int r = 0; r < n_ranks; ++r) {
      SvbHandle* h = m->ranks[(size_t)r];
      h->slab_lo = m->plan[(size_t)r].first; h->slab_hi = m->plan[(size_t)r].second;
      h->reach_lo = r > 0 ? m->plan[(size_t)r - 1].first : h->slab_lo;
      h->reach_hi = r + 1 < n_ranks ? m->plan[(size_t)r + 1].second : h->slab_hi;
    }
    return 0;
  }
  // ---- join the ranks: peer access, mailboxes (allocated for the largest slab), slab table
  rc = for_each_rank(m, [&](int r) -> int {
    SvbHandle* h = m->ranks[(size_t)r];
    CK(cudaSetDevice(h->device));
    for (int q = 0; q < n_ranks; ++q)
      if (q != r) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, h->device, m->devices[(size_t)q]));
        if (!can) return fail(h, SVB_COMM_ERROR, "device %d cannot access device %d's memory (peer access is required by svb_create_multi)", h->device, m->devices[(size_t)q]);
        const cudaError_t e = cudaDeviceEnablePeerAccess(m->devices[(size_t)q], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(h, SVB_CUDA_ERROR, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
      }
    h->rank = r; h->n_ranks = n_ranks;
    h->slab_lo = m->plan[(size_t)r].first; h->slab_hi = m->plan[(size_t)r].second;
    h->reach_lo = r > 0 ? m->plan[(size_t)r - 1].first : h->slab_lo;
    h->reach_hi = r + 1 < n_ranks ? m->plan[(size_t)r + 1].second : h->slab_hi;
    CK(h->comm_counts.ensure(256 + 64 + (size_t)n_ranks * 8));
    CK(cudaMemsetAsync(h->comm_counts.p, 0, 256 + 64 + (size_t)n_ranks * 8, h->stream));
    CK(cudaMallocHost(&h->h_counts, 64 + (size_t)n_ranks * 8));
    if (int rc_r = resize_particles(h, (size_t)h->n * 3 / 2 + 65536)) return rc_r;
    if (int rc_r = ensure_tile_capacity(h, h->tile_cap * 2 + 8192)) return rc_r;
    return mailbox_alloc(h, n_max);
  });
  if (rc) { multi_note_error(front, rc); return rc; }
  rc = for_each_rank(m, [&](int r) -> int {
    SvbHandle* h = m->ranks[(size_t)r];
    CK(cudaSetDevice(h->device));
    for (int q = 0; q < n_ranks; ++q) h->peer_mailbox[q] = q == r ? nullptr : m->ranks[(size_t)q]->mailbox.p;
    h->slabs = true;
    h->p2p = true;
    h->ipc_mapped = false;
    return mailbox_finish(h);
  });
  if (rc) multi_note_error(front, rc);
  return rc;
}

int multi_download(SvbHandle* front, SvbParticles* out) {
  SvbMulti* m = front->multi;
  out->n = m->n_global;
  std::vector<uint64_t> seen((size_t)m->ranks.size(), 0);
  int rc = for_each_rank(m, [&](int r) -> int {
    SvbHandle* h = m->ranks[(size_t)r];
    HostRows rows;
    rows.resize(h->n);
    rows.index.resize(h->n);
    SvbParticles view = rows.view();
    std::vector<uint64_t> orig(std::max<size_t>(h->n, 1));
    if (int rc_r = svb_download_resident(h, &view, orig.data())) return rc_r;
    const size_t got = view.n;
    auto put1 = [&](float* dst, const std::vector<float>& src) { if (dst) for (size_t q = 0; q < got; ++q) dst[orig[q]] = src[q]; };
    auto putk = [&](float* dst, const std::vector<float>& src, int k) {
      if (dst) for (size_t q = 0; q < got; ++q) std::memcpy(dst + orig[q] * k, src.data() + q * k, (size_t)k * 4);
    };
    for (size_t q = 0; q < got; ++q) {
      if (orig[q] >= m->n_global) return fail(h, SVB_COMM_ERROR, "a resident row carries original index %llu of %llu particles", (unsigned long long)orig[q], (unsigned long long)m->n_global);
      if (out->flags) out->flags[orig[q]] = rows.flags[q];
      if (out->collider_bits) out->collider_bits[orig[q]] = rows.bits[q];
    }
    put1(out->mass, rows.mass); put1(out->initial_volume, rows.vol); put1(out->mu_or_bulk_modulus, rows.p0); put1(out->lambda_or_exponent, rows.p1);
    put1(out->sand_alpha, rows.alpha); put1(out->viscosity_dynamic, rows.vd); put1(out->viscosity_bulk, rows.vb); put1(out->elastic_energies, rows.energy);
    putk(out->positions, rows.x, 3); putk(out->velocities, rows.v, 3); putk(out->position_gradients, rows.F, 9); putk(out->velocity_gradients, rows.C, 9);
    seen[(size_t)r] = got;
    return 0;
  });
  if (rc) { multi_note_error(front, rc); return rc; }
  uint64_t total = 0;
  for (uint64_t c : seen) total += c;
  if (total != m->n_global) return fail(front, SVB_COMM_ERROR, "the slab ranks hold %llu rows, the state has %llu particles", (unsigned long long)total, (unsigned long long)m->n_global);
  if (out->initial_positions && !front->initial_positions.empty()) std::memcpy(out->initial_positions, front->initial_positions.data(), front->initial_positions.size() * 4);
  return 0;
}

int multi_advance(SvbHandle* front, double target_time, float max_time_step, int32_t adaptive, const volatile int32_t* cancel, void (*progress)(void*, size_t), void* user) {
  SvbMulti* m = front->multi;
  const int rc = for_each_rank(m, [&](int r) -> int {
    return svb_advance(m->ranks[(size_t)r], target_time, max_time_step, adaptive, cancel, r == 0 ? progress : nullptr, user);
  });
  SvbHandle* h0 = m->ranks[0];
  front->time = h0->time;
  front->substeps = h0->substeps;
  front->adaptive = h0->adaptive;
  front->status = 0;
  front->launches = 0;
  front->last_advance_ms = 0;
  for (SvbHandle* r : m->ranks) {
    front->status |= r->status;
    front->launches += r->launches;
    front->last_advance_ms = std::max(front->last_advance_ms, r->last_advance_ms);
  }
  if (rc) multi_note_error(front, rc);
  return rc;
}

int multi_set_topology(SvbHandle* front, uint32_t n_colliders, const uint32_t* num_vertices, const uint32_t* num_triangles, const uint32_t* triangles) {
  SvbMulti* m = front->multi;
  const int rc = for_each_rank(m, [&](int r) -> int { return svb_set_topology(m->ranks[(size_t)r], n_colliders, num_vertices, num_triangles, triangles); });
  if (rc) multi_note_error(front, rc);
  return rc;
}
int multi_set_keyframes(SvbHandle* front, uint64_t frame, const SvbKeyframe* a, const SvbKeyframe* b) {
  SvbMulti* m = front->multi;
  const int rc = for_each_rank(m, [&](int r) -> int { return svb_set_keyframes(m->ranks[(size_t)r], frame, a, b); });
  if (rc) multi_note_error(front, rc);
  return rc;
}
void multi_set_option(SvbHandle* front, const char* name, double value) {
  for (SvbHandle* r : front->multi->ranks) svb_set_option(r, name, value);
}
void multi_destroy(SvbHandle* front) {
  for (SvbHandle* r : front->multi->ranks)
    if (r) svb_destroy(r);
  delete front->multi;
  front->multi = nullptr;
  delete front;
}
int multi_upload(SvbHandle* front, const SvbParticles* p, double time) {
  SvbMulti* m = front->multi;
  size_t old_max = 0;
  for (SvbHandle* r : m->ranks) old_max = std::max<size_t>(old_max, r->mb_mig_cap);
  if (p->n / m->ranks.size() / 4 + 65536 > old_max * 2)
    return fail(front, SVB_BAD_ARGUMENT, "svb_upload on a multi-device handle: the new state is much larger than the one the mailboxes were sized for; create a new handle");
  front->time = time;
  front->substeps = 0;
  front->status = 0;
  front->n = (uint32_t)p->n;
  front->initial_positions.assign((size_t)p->n * 3, 0.f);
  if (p->initial_positions && p->n) std::memcpy(front->initial_positions.data(), p->initial_positions, (size_t)p->n * 12);
  return multi_load(front, p, time, /*create=*/false);
}

}  // namespace

extern "C" int32_t svb_create_multi(const SvbConsts* consts, const SvbParticles* p, double time, const int32_t* devices, int32_t n_dev, SvbHandle** out) {
  if (!consts || !p || !out || !devices || n_dev < 1 || n_dev > SLAB_MAX_RANKS) return SVB_BAD_ARGUMENT;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SVB_CUDA_ERROR;  // no CPU fallback
  for (int k = 0; k < n_dev; ++k) {
    if (devices[k] < 0 || devices[k] >= count) return SVB_BAD_ARGUMENT;
    for (int q = 0; q < k; ++q)
      if (devices[q] == devices[k]) return SVB_BAD_ARGUMENT;
  }
  if (p->n > 0xfffffff0ull) return SVB_BAD_ARGUMENT;
  SvbHandle* front = new SvbHandle();
  *out = front;
  front->multi = new SvbMulti();
  front->multi->devices.assign(devices, devices + n_dev);
  front->multi->consts = *consts;
  front->consts = *consts;
  front->device = devices[0];
  front->time = time;
  front->n = (uint32_t)p->n;
  front->initial_positions.assign((size_t)p->n * 3, 0.f);
  if (p->initial_positions && p->n) std::memcpy(front->initial_positions.data(), p->initial_positions, (size_t)p->n * 12);
  return multi_load(front, p, time, /*create=*/true);
}
