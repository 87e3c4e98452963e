"""oracle/eigen_shim: the AVX2/FMA form of dot32 (what the reference arm runs) is bit-identical
to the scalar statement of the same order (what oracle/jpp_oracle.c and the CUDA kernels replay)."""
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]

SRC = r"""
#include <eigen3/Eigen/Core>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>
int main() {
#if !(defined(__AVX2__) && defined(__FMA__))
  std::printf("SKIP\n");
  return 0;
#endif
  std::mt19937 rng(123);
  std::uniform_real_distribution<float> d(-2.0f, 2.0f);
  long bad = 0, n = 0;
  for (int len = 0; len <= 300; ++len) {
    for (int rep = 0; rep < 40; ++rep) {
      std::vector<float> a(len + 3), b(len + 3);
      for (auto& x : a) x = d(rng) * (rep % 5 == 0 ? 1e-3f : 1.0f);
      for (auto& x : b) x = d(rng);
      const int off = rep % 3;  // unaligned starts
      float x = Eigen::shim::dot32(a.data() + off, b.data() + off, len);
      float y = Eigen::shim::dot32_scalar(a.data() + off, b.data() + off, len);
      bad += std::memcmp(&x, &y, 4) != 0;
      ++n;
    }
  }
  std::printf("%ld %ld\n", bad, n);
  return bad != 0;
}
"""


def test_avx2_dot32_is_bit_identical_to_the_scalar_order(tmp_path):
    src = tmp_path / "t.cc"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O3", "-march=x86-64-v3", "-ffp-contract=off", f"-I{ROOT / 'oracle' / 'eigen_shim'}",
                    str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    if r.stdout.strip() != "SKIP":
        bad, n = map(int, r.stdout.split())
        assert bad == 0 and n == 301 * 40
