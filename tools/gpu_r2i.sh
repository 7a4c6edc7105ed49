#!/bin/bash
TAG=${1:-r2i}
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
CS=$PWD/dummynode4graphlearning_b200/csrc
DN4GL_LIB=$CS/libdn4gl_pipetl.so timeout 300 python tools/pipe_timeline.py --dims 32 > gpurun_out/${TAG}_pipe_tl.jsonl 2> gpurun_out/${TAG}_pipe_tl.err
echo "pipe tl rc=$?"; tail -3 gpurun_out/${TAG}_pipe_tl.err
