#!/bin/bash
# sass_fp.sh <substring> -- md5 of the SASS instruction stream (addresses/encodings stripped) of every kernel in
# libu2pl_b200.so whose mangled name contains <substring>; used to prove a refactor left a validated kernel untouched.
LIB="$(dirname "$0")/../u2pl_b200/libu2pl_b200.so"
for f in $(cuobjdump -sass "$LIB" | grep "Function :" | awk '{print $3}' | grep "$1"); do
  n=$(cuobjdump -sass -fun "$f" "$LIB" 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's#/\*[0-9a-f]+\*/##g' | tee /tmp/sass_fp.$$ | wc -l)
  echo "$(md5sum < /tmp/sass_fp.$$ | cut -c1-12) $n $f"
done
rm -f /tmp/sass_fp.$$
