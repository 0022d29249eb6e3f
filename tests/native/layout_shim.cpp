// Host build of the product's particle layout (squishy_volumes_b200/csrc/svb_device.cuh: Field, ParticleBuf) for tests/test_layout.py.
#include <cstring>
#include "../../squishy_volumes_b200/csrc/svb_device.cuh"

extern "C" {
int shim_layout_counts(int* nfields, int* nquads, int* nwords) {
  *nfields = svb::NFIELDS; *nquads = svb::NQUADS; *nwords = svb::NWORDS;
  return 0;
}
// offset (in 4-byte words from the buffer base) of word `field` of particle `i` in a buffer of capacity `cap`, through the float and the u32 accessor
long long shim_word_offset(int field, unsigned long long cap, unsigned long long i, int as_u32) {
  static uint32_t dummy;
  svb::ParticleBuf P{&dummy, (size_t)cap};
  if (as_u32) return (long long)(&P.u(field)[i] - &dummy);
  return (long long)(&P.f(field)[i] - reinterpret_cast<float*>(&dummy));
}
// offset (in words) of quad `q` of particle `i`
long long shim_quad_offset(int q, unsigned long long cap, unsigned long long i) {
  static uint32_t dummy;
  svb::ParticleBuf P{&dummy, (size_t)cap};
  return (long long)(reinterpret_cast<uint32_t*>(P.q(q) + i) - &dummy);
}
int shim_field_ids(int* out) {   // PX, PFLAGS, PF, PMASS, PVOL, PP0, PP1, PALPHA, PVD, PVB, PBITS, PORIG, PV, PC
  const int ids[] = {svb::PX, svb::PFLAGS, svb::PF, svb::PMASS, svb::PVOL, svb::PP0, svb::PP1, svb::PALPHA, svb::PVD, svb::PVB, svb::PBITS, svb::PORIG, svb::PV, svb::PC};
  for (int k = 0; k < 14; ++k) out[k] = ids[k];
  return 14;
}
}
