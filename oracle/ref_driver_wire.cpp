// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED md-flexible particle (de)serialisation
// (examples/md-flexible/src/ParticleSerializationTools.cpp, compiled where it lies, single-site mode): the MPI wire
// format that apb_serialize_particles / apb_deserialize_particles interoperate with. Own library
// (oracle/_ref/libautopas_ref_wire.so) because TypeDefinitions.h needs the reference's MD_FLEXIBLE_MODE macros.
#include <cstdint>
#include <cstring>
#include <vector>

#include "ParticleSerializationTools.h"

extern "C" {
// cols: 12 arrays (x y z vx vy vz fx fy fz oldFx oldFy oldFz); out: 120 bytes per particle. Returns bytes written.
int64_t ref_wire_serialize(int64_t n, const double *const *cols, const int64_t *id, const int64_t *type, const int64_t *own,
                           char *out) {
  std::vector<char> bytes;
  for (int64_t i = 0; i < n; ++i) {
    ParticleType p({cols[0][i], cols[1][i], cols[2][i]}, {cols[3][i], cols[4][i], cols[5][i]}, static_cast<unsigned long>(id[i]),
                   static_cast<unsigned long>(type[i]));
    p.setF({cols[6][i], cols[7][i], cols[8][i]});
    p.setOldF({cols[9][i], cols[10][i], cols[11][i]});
    p.setOwnershipState(static_cast<autopas::OwnershipState>(own[i]));
    ParticleSerializationTools::serializeParticle(p, bytes);
  }
  std::memcpy(out, bytes.data(), bytes.size());
  return static_cast<int64_t>(bytes.size());
}
// the inverse: fills the 12 columns, ids, types, ownership states from n records; returns the number of particles
int64_t ref_wire_deserialize(int64_t numBytes, const char *in, double *const *cols, int64_t *id, int64_t *type, int64_t *own) {
  std::vector<char> bytes(in, in + numBytes);
  std::vector<ParticleType> ps;
  ParticleSerializationTools::deserializeParticles(bytes, ps);
  for (size_t i = 0; i < ps.size(); ++i) {
    const auto &p = ps[i];
    for (int d = 0; d < 3; ++d) {
      cols[d][i] = p.getR()[d];
      cols[3 + d][i] = p.getV()[d];
      cols[6 + d][i] = p.getF()[d];
      cols[9 + d][i] = p.getOldF()[d];
    }
    id[i] = static_cast<int64_t>(p.getID());
    type[i] = static_cast<int64_t>(p.getTypeId());
    own[i] = static_cast<int64_t>(p.getOwnershipState());
  }
  return static_cast<int64_t>(ps.size());
}
}
