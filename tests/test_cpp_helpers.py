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
