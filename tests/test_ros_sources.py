"""the files a ROS maintainer builds -- node main, nodelet class, plugin description, catkin
CMakeLists -- exist, keep the reference's names, and compile (syntax only: this image has no ROS) against
declaration-only stand-ins of the ROS headers (tests/ros_stubs) AND the real facade headers."""
import os
import re
import shutil
import subprocess
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROS = os.path.join(ROOT, "realtime_urdf_filter_b200", "host", "ros")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
@pytest.mark.parametrize("src", ["realtime_urdf_filter_node.cpp", "realtime_urdf_filter_nodelet.cpp"])
def test_ros_sources_compile_against_stub_headers(src):
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
                          "-I", os.path.join(ROOT, "tests", "ros_stubs"), "-I", os.path.join(ROOT, "include"),
                          os.path.join(ROS, src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_plugin_description_keeps_the_reference_names():
    # plugins/nodelet_plugins.xml:1-2 of the reference: library path and class name are what launch files load
    lib = ET.parse(os.path.join(ROS, "nodelet_plugins.xml")).getroot()
    assert lib.tag == "library" and lib.get("path") == "lib/librealtime_urdf_filter_nodelet"
    cls = lib.find("class")
    assert cls.get("name") == "realtime_urdf_filter/RealtimeURDFFilterNodelet"
    assert cls.get("type") == "realtime_urdf_filter::RealtimeURDFFilterNodelet"
    assert cls.get("base_class_type") == "nodelet::Nodelet"
    src = open(os.path.join(ROS, "realtime_urdf_filter_nodelet.cpp")).read()
    assert "PLUGINLIB_EXPORT_CLASS(realtime_urdf_filter::RealtimeURDFFilterNodelet, nodelet::Nodelet)" in src
    pkg = ET.parse(os.path.join(ROS, "package.xml")).getroot()
    assert pkg.find("name").text == "realtime_urdf_filter"
    assert pkg.find("export/nodelet").get("plugin") == "${prefix}/nodelet_plugins.xml"


def test_cmake_targets_and_node_behaviour_follow_the_reference():
    cm = open(os.path.join(ROS, "CMakeLists.txt")).read()
    assert re.search(r"project\(realtime_urdf_filter\b", cm)
    for target in ("add_library(urdf_filter ", "add_executable(realtime_urdf_filter ", "add_library(realtime_urdf_filter_nodelet "):
        assert target in cm, target
    assert "100a" in cm and "-fmad=false" in cm and "fast_math" not in cm.replace("never -use_fast_math", "")
    for s in re.findall(r"\$\{RUF_PKG\}/(\S+?\.(?:cu|cpp))", cm):          # every listed source exists
        assert os.path.exists(os.path.join(ROOT, "realtime_urdf_filter_b200", s)), s
    node = open(os.path.join(ROS, "realtime_urdf_filter_node.cpp")).read()
    # src/realtime_urdf_filter.cpp:39-50: node name, private handle, spin under a runtime_error catch
    assert 'ros::init(argc, argv, "realtime_urdf_filter")' in node and 'ros::NodeHandle nh("~")' in node
    assert re.search(r"try\s*{\s*ros::spin\(\);\s*}\s*catch \(const std::runtime_error &e\)", node)
    bridge = open(os.path.join(ROS, "ros_bridge.h")).read()
    for name in ('"input_depth"', '"output_depth"', '"output_mask"', '"fixed_frame"', '"camera_frame"', '"camera_offset"',
                 '"depth_distance_threshold"', '"filter_replace_value"', '"show_gui"', '"models"', '"tf_prefix"',
                 '"geometry_type"', '"scale"', '"ignore"'):
        assert name in bridge, name
