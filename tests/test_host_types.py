"""Host-side mirror types (roargraph_b200/host/include/efanna2e) against the reference's own headers: compiled side by
side from the same driver, traces must be identical.  Needs /root/reference (this container); skipped on the GPU box."""
import os
import subprocess

import pytest

from conftest import ROOT

REF_INC = "/root/reference/include"
MIRROR_INC = os.path.join(ROOT, "roargraph_b200", "host", "include")
DRIVER = os.path.join(ROOT, "tests", "cpp", "npq_driver.cpp")


def _build(tmp_path, name, inc):
    exe = str(tmp_path / name)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I", inc, DRIVER, "-o", exe])
    return exe


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_INC, "efanna2e", "neighbor.h")), reason="/root/reference not mounted")
def test_neighbor_priority_queue_mirror_traces_equal_reference(tmp_path):
    ref = _build(tmp_path, "npq_ref", REF_INC)
    ours = _build(tmp_path, "npq_ours", MIRROR_INC)
    for cap, n_ops, id_range, seed in ((1, 300, 8, 1), (4, 2000, 16, 2), (16, 5000, 64, 3), (100, 20000, 4000, 4),
                                       (7, 3000, 7, 5)):
        args = [str(cap), str(n_ops), str(id_range), str(seed)]
        a = subprocess.run([ref, *args], capture_output=True, text=True, check=True).stdout
        b = subprocess.run([ours, *args], capture_output=True, text=True, check=True).stdout
        assert a == b and a.count("\n") > 3, (cap, n_ops, id_range, seed)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_INC, "efanna2e", "util.h")), reason="/root/reference not mounted")
def test_normalize_mirror_is_bit_identical_to_the_compiled_reference(tmp_path):
    """efanna2e::normalize as the reference's release flags compile it (8 lane sums, RSQRTSS + one Newton step, multiply)
    vs the mirror's restatement of that instruction sequence, dim % 8 == 0 (the reference's own precondition for cosine,
    SURVEY.md 9.10): identical bits.  This is what makes the cosine CLI path feed the device the reference's exact rows."""
    driver = os.path.join(ROOT, "tests", "cpp", "normalize_driver.cpp")
    ref, ours = str(tmp_path / "norm_ref"), str(tmp_path / "norm_ours")
    # the reference's flags (CMakeLists.txt:28) / the mirror's (roargraph_b200/host/Makefile)
    subprocess.check_call(["g++", "-std=c++17", "-Ofast", "-DNDEBUG", "-march=native", "-mtune=native", "-ftree-vectorize",
                           "-w", "-I", REF_INC, driver, "-o", ref])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-w", "-I", MIRROR_INC, driver, "-o", ours])
    for dim, rows, seed in ((8, 50, 1), (64, 200, 2), (200, 300, 3), (512, 100, 4), (104, 64, 5)):
        a = subprocess.run([ref, str(dim), str(rows), str(seed)], capture_output=True, text=True, check=True).stdout
        b = subprocess.run([ours, str(dim), str(rows), str(seed)], capture_output=True, text=True, check=True).stdout
        assert a == b and len(a) > dim * rows * 8, (dim, rows, seed)


def test_normalize_rows_is_thread_count_independent(tmp_path):
    src = tmp_path / "nr.cpp"
    src.write_text(r"""
#include <cstdio>
#include <cstring>
#include <vector>
#include "efanna2e/util.h"
int main() {
  const size_t n = 50000, dim = 40;
  std::vector<float> a(n * dim), b;
  for (size_t i = 0; i < a.size(); ++i) a[i] = (float)((i * 2654435761u) % 1000) / 37.0f - 13.0f;
  b = a;
  efanna2e::normalize_rows(a.data(), n, dim, 1);
  efanna2e::normalize_rows(b.data(), n, dim, 7);
  std::vector<float> c(a.size());
  for (size_t i = 0; i < a.size(); ++i) c[i] = (float)((i * 2654435761u) % 1000) / 37.0f - 13.0f;
  for (size_t r = 0; r < n; ++r) efanna2e::normalize<float>(c.data() + r * dim, dim);
  printf("%d %d\n", memcmp(a.data(), b.data(), a.size() * 4), memcmp(a.data(), c.data(), a.size() * 4));
  return 0;
}
""")
    exe = str(tmp_path / "nr")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I", MIRROR_INC, str(src), "-o", exe])
    assert subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split() == ["0", "0"]


def test_mirror_headers_are_self_contained(tmp_path):
    """Every public header of the mirror compiles on its own (what a maintainer's include-path switch relies on)."""
    for h in ("efanna2e/distance.h", "efanna2e/neighbor.h", "efanna2e/parameters.h", "efanna2e/util.h", "efanna2e/index.h",
              "efanna2e/exceptions.h", "index_bipartite.h"):
        src = tmp_path / "t.cpp"
        src.write_text(f'#include "{h}"\nint main() {{ return 0; }}\n')
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", MIRROR_INC,
                               "-I", os.path.join(ROOT, "include"), str(src)])


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_rendezvous_gives_every_worker_the_same_verdict(tmp_path, n):
    """roargraph_b200/csrc/rg_rendezvous.h (the agreement the rg_group_* workers reach before they enter a collective): every
    thread sees the AND of all reports, round after round, with the object reused."""
    exe = str(tmp_path / "rendezvous")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "roargraph_b200", "csrc"),
                           os.path.join(ROOT, "tests", "cpp", "rendezvous_driver.cpp"), "-o", exe])
    for fail_every in (1, 3, 1000):
        out = subprocess.run([exe, str(n), "3000", str(fail_every)], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0 and out.stdout.startswith("ok 3000 "), out.stdout + out.stderr
