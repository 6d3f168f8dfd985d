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
    // the N4 curves through the same template: sizes follow the curve (gmsm_affine_bytes / gmsm_scalar_bytes), [1]G = G
    static_assert(sizeof(secp256k1::G1Affine) == 64 && sizeof(secp256k1::G1::Scalar) == 32, "secp256k1 layout");
    static_assert(sizeof(bw6761::G1Affine) == 192 && sizeof(bw6761::G2Affine) == 192 && sizeof(bw6761::G1::Scalar) == 48, "bw6-761 layout");
    static_assert(sizeof(bls24315::G1Affine) == 80 && sizeof(bls24317::G1Jac) == 120, "bls24 layout");
    static_assert(sizeof(bw6633::G1Affine) == 160 && sizeof(bw6633::G1::Scalar) == 40, "bw6-633 layout");
    CHECK(gmsm_affine_bytes(GMSM_BW6761_G1) == sizeof(bw6761::G1Affine) && gmsm_scalar_bytes(GMSM_BW6761_G1) == sizeof(bw6761::G1::Scalar));
    CHECK(gmsm_affine_bytes(GMSM_BW6633_G2) == sizeof(bw6633::G2Affine) && gmsm_scalar_bytes(GMSM_BW6633_G2) == sizeof(bw6633::G2::Scalar));
    {
      secp256k1::G1Affine S;   // ecc/secp256k1/secp256k1.go:51-52, Montgomery limbs
      S.X = {0xd7362e5a487e2097ull, 0x231e295329bc66dbull, 0x979f48c033fd129cull, 0x9981e643e9089f48ull};
      S.Y = {0xb15ea6d2d3dbabe2ull, 0x8dfc5d5d1f1dc64dull, 0x70b6b59aac19c136ull, 0xcf3f851fd4a582d6ull};
      const secp256k1::G1::Scalar s_one = {0x402da1732fc9bebfull, 0x4551231950b75fc4ull, 0x0000000000000001ull, 0x0000000000000000ull};   // 2^256 mod r
      CHECK(secp256k1::G1::MultiExpAffine({S}, {s_one}) == S);
      CHECK(secp256k1::G1::MultiExpAffine({S, S}, {s_one, {0, 0, 0, 0}}) == S);
      bw6761::G1Affine B;      // ecc/bw6-761/bw6-761.go:97-98
      B.X = {0xd6e42d7614c2d770ull, 0x4bb886eddbc3fc21ull, 0x64648b044098b4d2ull, 0x1a585c895a422985ull, 0xf1a9ac17cf8685c9ull, 0x352785830727aea5ull, 0xddf8cb12306266feull, 0x6913b4bfbc9e949aull, 0x3a4b78d67ba5f6abull, 0x0f481c06a8d02a04ull, 0x91d4e7365c43edacull, 0x00f4d17cd48beca5ull};
      B.Y = {0x97e805c4bd16411full, 0x870d844e1ee6dd08ull, 0x1eba7a37cb9eab4dull, 0xd544c4df10b9889aull, 0x8fe37f21a33897beull, 0xe9bf99a43a0885d2ull, 0xd7ee0c9e273de139ull, 0xaa6a9ec7a38dd791ull, 0x8f95d3fcf765da8eull, 0x42326e7db7357c99ull, 0xe217e407e218695full, 0x009d1eb23b7cf684ull};
      const bw6761::G1::Scalar b_one = {0x02cdffffffffff68ull, 0x51409f837fffffb1ull, 0x9f7db3a98a7d3ff2ull, 0x7b4e97b76e7c6305ull, 0x4cf495bf803c84e8ull, 0x008d6661e2fdf49aull};      // 2^384 mod r: six words
      CHECK(bw6761::G1::MultiExpAffine({B}, {b_one}) == B);
      auto bb = bw6761::G1::BatchScalarMultiplication(B, {b_one, {0, 0, 0, 0, 0, 0}});
      CHECK(bb[0] == B && bb[1].IsInfinity());
      bw6633::G2Affine W;      // ecc/bw6-633/bw6-633.go:92-93
      W.X = {0x6ee4df974ed3c43full, 0x1e1809e474afe2ddull, 0xb34aa235e12dde2cull, 0x4540cf8e03270cedull, 0x2940ff4f76a4d4a1ull, 0x8ab5078bb3ba80adull, 0xbaa6212266c6d100ull, 0x857ea2365030d3e5ull, 0x9d0a7655931c6a61ull, 0x00c23b4dbea27983ull};
      W.Y = {0xbae3a0f5dcf24f3cull, 0x0f7004d8c9857c8eull, 0x6b873314adff62e9ull, 0x6573f1345f595ee2ull, 0x3de8706a06fc2e9cull, 0xa661298f9de21b6cull, 0x6e56bf005bed0261ull, 0x0d73e1cdd03bed1cull, 0xe5e617018b191c3aull, 0x00f56b107a073952ull};
      const bw6633::G2::Scalar w_one = {0xd4f76127b60fffcbull, 0x4f9a69ccdeaf967eull, 0xe1dfea7c5cb86f92ull, 0x1dcdc608a9406596ull, 0x03c9fd706b15a144ull};      // 2^320 mod r: five words
      CHECK(bw6633::G2::MultiExpAffine({W}, {w_one}) == W);
    }
  }
  std::printf(fails ? "MIRROR_FAILED\n" : "MIRROR_OK\n");
  return fails ? 1 : 0;
}
