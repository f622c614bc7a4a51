"""Host-only unit tests of C++ pieces of the product library that need no GPU (compiled on the fly with g++)."""
import os
import subprocess

from conftest import ROOT


def test_header_is_valid_c_and_cxx(tmp_path):
    """include/meepo.h is the drop-in boundary: it must compile as plain C99 and as C++ on its own."""
    src_c = tmp_path / "use.c"
    src_c.write_text('#include "meepo.h"\nint main(void) { meepo_config c; (void)c; return (int)meepo_owner(1u, 2u) * 0; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_c)])
    src_cc = tmp_path / "use.cc"
    src_cc.write_text('#include "meepo.h"\nint main() { meepo_stats_t s{}; return (int)s.size; }\n')
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_cc)])


def test_c_example_compiles_and_links(tmp_path):
    """examples/minimal.c against include/meepo.h and libmeepo.so (linking needs no GPU; running does)."""
    lib_dir = os.path.join(ROOT, "meepoembedding_b200")
    if not os.path.exists(os.path.join(lib_dir, "libmeepo.so")):
        import __graft_entry__

        __graft_entry__.build()
    exe = tmp_path / "minimal"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "minimal.c"), "-L", lib_dir, "-lmeepo",
                           f"-Wl,-rpath,{lib_dir}", "-o", str(exe)])
    assert exe.exists()
