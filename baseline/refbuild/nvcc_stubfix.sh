#!/bin/bash
# Usage: nvcc_stubfix.sh <nvcc command line ... -c -o OUT.o FILE.cu>
# Three .cu files of the unmodified reference (leaky_relu, dropout, quantized_conv) instantiate a CUB scan whose functor type is
# cuda::std::plus<void>.  nvcc's generated host stub (NOT reference source) then contains `typedef cuda::std::...` at global scope, where
# the reference's global `using namespace mshadow;` (src/operator/linalg.h:32) makes `cuda` ambiguous (mshadow::cuda vs libcu++'s ::cuda).
# Fix on the generated intermediate only: keep nvcc's intermediates, qualify that typedef with `::`, re-run nvcc's own final host compile.
set -e
K=$(mktemp -d /tmp/stubfix.XXXX)
"$@" --keep --keep-dir "$K" > "$K/first.log" 2>&1 && { rm -rf "$K"; exit 0; }
sed -i 's/^typedef cuda::std::/typedef ::cuda::std::/' "$K"/*.cudafe1.stub.c
LAST=$("$@" --keep --keep-dir "$K" --dryrun 2>&1 | sed 's/^#\$ //' | grep -E '^g\+\+ .* -c -x c\+\+' | tail -1)
eval "$LAST"
rm -rf "$K"
