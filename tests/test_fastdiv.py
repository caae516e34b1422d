"""csrc/fastdiv.h (the division-free tile decode of the persistent tcgen05 kernels): q = umulhi(n, mul) >> shr must equal n / d for
every dividend the kernels can form (0 <= n < 2^31) -- compiled for the host and checked exhaustively over small divisors and
dividends, on multiples of d and their neighbours, and at the top of the range."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include <stdio.h>
#include <stdlib.h>
#include "fastdiv.h"
static long bad = 0;
static void chk(int n, const FastDiv& f) { if (n >= 0 && fast_div(n, f) != n / f.d) { if (++bad < 5) printf("bad %d / %d\n", n, f.d); } }
int main() {
  for (int d = 1; d <= 40000; ++d) {
    const FastDiv f = make_fastdiv(d);
    for (int n = 0; n < 600; ++n) chk(n, f);
    for (long q = 1; q < 64; ++q) { const long n = q * d; if (n < 2147483647L) { chk((int)n - 1, f); chk((int)n, f); chk((int)n + 1, f); } }
    for (int k = 0; k < 64; ++k) chk(2147483647 - k, f);
    const long top = (2147483647L / d) * d;            // largest multiple of d in range
    chk((int)top, f); chk((int)top - 1, f);
  }
  const int big[] = {65535, 65536, 65537, 1000003, 1 << 20, (1 << 20) + 1, 1 << 30, (1 << 30) + 7, 2147483646, 2147483647};
  unsigned long long s = 88172645463325252ull;
  for (int d : big) {
    const FastDiv f = make_fastdiv(d);
    for (int k = 0; k < 400000; ++k) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; chk((int)(s & 0x7fffffff), f); }
    for (int k = 0; k < 64; ++k) { chk(2147483647 - k, f); chk(k, f); }
  }
  printf("bad=%ld\n", bad);
  return bad != 0;
}
'''


def test_fast_div_is_exact(tmp_path):
    src = tmp_path / "fd.cpp"
    src.write_text(SRC)
    exe = tmp_path / "fd"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "biggan_tensorflow_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout
