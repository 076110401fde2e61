# builds experiment variants of the library into variants/ (git-ignored): tools/build_variants.sh NAME "-DFOO=1 -DBAR=2" ...
set -e
cd "$(dirname "$0")/../soap_b200/csrc"
mkdir -p ../../variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  rm -rf /tmp/var_$name && mkdir -p /tmp/var_$name
  for f in mesh scan chunk halos moments tier proj iter seq; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-O2 $defs -c $f.cu -o /tmp/var_$name/$f.o 2>/dev/null &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so /tmp/var_$name/*.o
  echo built $name
done
