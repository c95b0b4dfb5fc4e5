#!/bin/bash
# A/B builds of libb200pt.so with extra -D flags: tools/build_variants.sh name "-DPT_OPT_X ..." [name2 "..."] ...
# -> learn_path_tracing_b200/variants/libb200pt_<name>.so (travels to the GPU box; select with PT_LIB_PATH)
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/learn_path_tracing_b200/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  tmp=/tmp/ptvar_$name; rm -rf $tmp; mkdir -p $tmp/learn_path_tracing_b200 $tmp/include
  cp -r $root/learn_path_tracing_b200/csrc $tmp/learn_path_tracing_b200/csrc; cp $root/include/*.h $tmp/include/
  rm -rf $tmp/learn_path_tracing_b200/csrc/*.o $tmp/learn_path_tracing_b200/csrc/obj $tmp/learn_path_tracing_b200/csrc/obj_exp
  make -s -C $tmp/learn_path_tracing_b200/csrc -j8 EXTRA="$flags" OUT=$root/learn_path_tracing_b200/variants/libb200pt_$name.so &
done
wait
ls -la $root/learn_path_tracing_b200/variants/
