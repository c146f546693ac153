# developer tool: build a variant of the library with extra nvcc flags for the named sources (others reuse build/obj) -> variants/lib_<name>.so
# usage: bash tools/build_variant.sh <name> "<flags>" file1.cu [file2.cu ...]     (A/B on one GPU box: copy the variant over cruse_b200/libcruse_sm100.so)
set -e
name=$1; flags=$2; shift 2
python -m cruse_b200.build > /dev/null
mkdir -p variants/obj_$name
objs=""
for o in build/obj/*.o; do
  b=$(basename $o .o); use=$o
  for f in "$@"; do
    if [ "$b.cu" = "$f" ]; then
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -I include $flags -c cruse_b200/csrc/$f -o variants/obj_$name/$b.o
      use=variants/obj_$name/$b.o
    fi
  done
  objs="$objs $use"
done
nvcc -shared -o variants/lib_$name.so $objs
ls -la variants/lib_$name.so
