#!/bin/bash
# Produces tests/golden/ref_downstream.npz = outputs of the REFERENCE's own MapPointNormal / n_scan_normal_reg / RSCManager on the
# committed synthetic scans.  Needs what this repository's build container does not have: the reference's docker image
# (tbv_slam/docker/Dockerfile: ros:noetic-perception + Ceres 2.1.0) with an unmodified checkout of dan11003/tbv_slam_public built in a
# catkin workspace.  Run ON A MACHINE THAT HAS THEM:
#
#   docker build -t tbv_ref <reference>/tbv_slam/docker
#   docker run --rm -v <reference>:/ws/src/tbv_slam_public -v $PWD:/repo tbv_ref bash /repo/tools/make_ref_fixtures.sh
#
# Steps: (1) write the input scans (python, numpy only); (2) catkin build of the reference (its own build system, in ITS tree — nothing of
# it is copied here) + the dumper as one extra executable linked against its libraries; (3) run the dumper; (4) convert its text output.
# Afterwards `pytest tests/test_ref_downstream_cpu.py` (oracle vs reference) and `-m gpu tests/test_ref_downstream_gpu.py` (CUDA path vs
# reference) stop skipping, and DESIGN.md §2's "parity unpinned" for rows a5-a20 can be turned into "pinned".
set -euo pipefail
REPO=${REPO:-/repo}
WS=${WS:-/ws}
cd "$REPO"
python3 tools/ref_fixtures/ref_downstream_io.py write-input /tmp/ref_scans.bin
source /opt/ros/noetic/setup.bash
cd "$WS"
catkin build cfear_radarodometry place_recognition_radar -DCMAKE_BUILD_TYPE=Release
source devel/setup.bash
INC="-I$WS/src/tbv_slam_public/cfear_radarodometry/include -I$WS/src/tbv_slam_public/place_recognition_radar/include $(pkg-config --cflags eigen3 opencv4) -I/usr/include/pcl-1.10 -I/opt/ros/noetic/include"
LIBS="-L$WS/devel/lib -lcfear_radarodometry -lplace_recognition_radar -L/opt/ros/noetic/lib -lcv_bridge -lroscpp -lrostime -lrosconsole -lroscpp_serialization $(pkg-config --libs opencv4) -lpcl_common -lpcl_kdtree -lpcl_search -lpcl_filters -lceres -lglog -lboost_system -lboost_serialization"
g++ -std=c++14 -O3 $INC -o /tmp/dump_ref_downstream "$REPO/tools/ref_fixtures/dump_ref_downstream.cpp" $LIBS -Wl,-rpath,$WS/devel/lib:/opt/ros/noetic/lib
/tmp/dump_ref_downstream /tmp/ref_scans.bin /tmp/ref_downstream.txt
cd "$REPO"
python3 tools/ref_fixtures/ref_downstream_io.py convert /tmp/ref_downstream.txt tests/golden/ref_downstream.npz
echo "wrote tests/golden/ref_downstream.npz"
