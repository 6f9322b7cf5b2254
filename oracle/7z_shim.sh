#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Stand-in for the external `7z` executable, which is not installed in
# this image.  The reference shells out to it (compression.cpp:308, decompression.cpp:34); the 7z
# stage is outside the hot path and untimed, so a copy keeps both reference programs runnable.
#   7z a -mx=9 OUT.7z IN      -> cp IN OUT.7z
#   7z e ARC -oDIR -y         -> cp ARC DIR/<basename of ARC without its last extension>
case "$1" in
  a)
    cp -- "$4" "$3" ;;
  e)
    arc="$2"
    dir="${3#-o}"
    base=$(basename -- "$arc")
    stem="${base%.*}"
    mkdir -p -- "$dir"
    # extracting onto itself (archive already sits in DIR under its stem name) is a no-op
    if [ "$(readlink -f -- "$arc")" != "$(readlink -f -- "$dir/$stem")" ]; then
      cp -- "$arc" "$dir/$stem"
    fi ;;
  *)
    echo "7z shim: unsupported command $1" >&2; exit 2 ;;
esac
