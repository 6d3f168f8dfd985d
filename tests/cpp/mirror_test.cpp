// Drives the C++ host mirror (include/gmsm.hpp) the way the reference's tests drive MultiExp.
//   argv[1] == "nogpu": only the host-side error behaviour + the loud refusal without a device
//   otherwise         : tiny known answers on the GPU: [1]G = G, [1]G + [1]G = [2]G, zero scalars -> infinity
#include <cstdio>
#include <cstring>
#include <string>

#include "gmsm.hpp"

using namespace gmsm_host;

static int fails = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAIL line %d: %s\n", __LINE__, #cond); fails++; } } while (0)

int main(int argc, char** argv) {
  const bool nogpu = argc > 1 && std::string(argv[1]) == "nogpu";
  // bn254 G1 generator (1, 2) and [2]G, fr One -- Montgomery limbs (ecc/bn254/bn254.go:111-113)
  bn254::G1Affine G, G2x;
  G.X = {0xd35d438dc58f0d9dull, 0x0a78eb28f5c70b3dull, 0x666ea36f7879462cull, 0x0e0a77c19a07df2full};
  G.Y = {0xa6ba871b8b1e1b3aull, 0x14f1d651eb8e167bull, 0xccdd46def0f28c58ull, 0x1c14ef83340fbe5eull};
  G2x.X = {0xe10460b6c3e7ea38ull, 0xbc0b548b438e5469ull, 0xc2822db40c0ac2ecull, 0x13227397098d014dull};
  G2x.Y = {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x04644e72e131a029ull};
  const bn254::G1::Scalar one = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  const bn254::G1::Scalar zero = {0, 0, 0, 0};

  // error behaviour of the reference (multiexp.go:61-71)
  try { bn254::G1Jac().MultiExp({G, G}, {one}); CHECK(false); }
  catch (const Error& e) { CHECK(std::string(e.what()) == "len(points) != len(scalars)"); }
  try { bn254::G1Jac().MultiExp({G}, {one}, MultiExpConfig{1025}); CHECK(false); }
  catch (const Error& e) { CHECK(std::string(e.what()) == "invalid config: config.NbTasks > 1024"); }

  if (nogpu) {
    try { bn254::G1Jac().MultiExp({G}, {one}); CHECK(false); }   // no CPU fallback: must refuse loudly
    catch (const Error& e) { CHECK(std::string(e.what()).find("no CUDA device") != std::string::npos); }
  } else {
    CHECK(bn254::G1::MultiExpAffine({G}, {one}) == G);
    CHECK(bn254::G1::MultiExpAffine({G, G}, {one, one}) == G2x);
    CHECK(bn254::G1::MultiExpAffine({G, G2x}, {zero, zero}).IsInfinity());
    bn254::G1Jac j;
    j.MultiExp({}, {});
    CHECK(j.IsInfinity());
    auto bs = bn254::G1::BatchScalarMultiplication(G, {zero, one});
    CHECK(bs[0].IsInfinity() && bs[1] == G);
    // resident bases, before and after the window tables: G + G = [2]G, sub-range, reference error string
    bn254::G1::ResidentBases rb({G, G, G2x});
    for (int pass = 0; pass < 2; pass++) {
      auto r2 = rb.MultiExp({one, one});
      CHECK(r2.X == G2x.X && r2.Y == G2x.Y && !r2.IsInfinity());
      auto r1 = rb.MultiExp({one}, {}, 2);
      CHECK(r1.X == G2x.X && r1.Y == G2x.Y);
      try { rb.MultiExp({one, one}, {}, 2); CHECK(false); }
      catch (const Error& e) { CHECK(std::string(e.what()) == "len(points) != len(scalars)"); }
      if (pass == 0) CHECK(rb.Precompute(5) == 5);
    }
  }
  std::printf(fails ? "MIRROR_FAILED\n" : "MIRROR_OK\n");
  return fails ? 1 : 0;
}
