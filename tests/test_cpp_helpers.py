"""CPU: the C++ facade's node-cuboid helper templates (creat_all_nodes / for_each_nodes / delete_all_nodes,
model_grid_map.hpp:90-138) compile against the C-ABI library and keep the reference's contract."""
import os
import subprocess

from conftest import ROOT


def test_cuboid_helper_templates(tmp_path):
    from welding_robot_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    exe = str(tmp_path / "helpers_test")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include", "welding_robot_b200"),
                    os.path.join(ROOT, "tests", "cpp", "helpers_test.cpp"), "-o", exe, "-L" + libdir, "-lwrgpu",
                    "-Wl,-rpath," + libdir, "-ldl", "-lpthread", "-lrt"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == "ok", r.stdout + r.stderr


def test_cpp_bs_basic_setparam_matches_reference_fixture(tmp_path):
    """The C++ BS_Basic facade (include/welding_robot_b200/compat/BSplineBasic.h): SetParam's knots and control points, bit for bit
    against the fixture of the unmodified reference header, for the demo's two instantiations and a mixed-constraint one."""
    import sys
    import numpy as np
    from conftest import GOLDEN
    from welding_robot_b200 import _lib
    sys.path.insert(0, GOLDEN)
    import make_golden as MG
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    exe = str(tmp_path / "bspline_test")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include", "welding_robot_b200"),
                    os.path.join(ROOT, "tests", "cpp", "bspline_test.cpp"), "-o", exe, "-L" + libdir, "-lwrgpu",
                    "-Wl,-rpath," + libdir, "-ldl", "-lpthread", "-lrt"], check=True)
    fx = np.load(os.path.join(GOLDEN, "ref_bspline.npz"))
    hexs = lambda a: " ".join("%08x" % v for v in np.ascontiguousarray(a, np.float32).ravel().view(np.uint32))   # noqa: E731
    done = 0
    for i, case in enumerate(MG.BSPLINE_CASES):
        if case[:3] not in ((0, 0, 0), (2, 2, 2), (3, 2, 1)):
            continue
        init, fin, mid, tf, _, _ = MG.bspline_inputs(case)
        text = "%d %d %d %d %s\n%s\n%s\n%s\n" % (case[0], case[1], case[2], mid.shape[0], hexs(np.float32(tf)), hexs(init), hexs(fin), hexs(mid[:, :3]))
        r = subprocess.run([exe], input=text, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        knots, cps = r.stdout.strip().split("\n")
        assert knots.split() == hexs(fx["knots%d" % i]).split(), case
        assert cps.split() == hexs(fx["cps%d" % i]).split(), case
        done += 1
    assert done == 3
