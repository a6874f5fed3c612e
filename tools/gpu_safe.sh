# source me: run "<seconds> <label> <cmd...>" with a hard timeout; the FIRST hang (rc 124/137) aborts the whole script so a
# deadlocked kernel cannot burn the GPU budget command after command
run() {
  local t=$1 label=$2; shift 2
  timeout -k 5 "$t" "$@"
  local rc=$?
  echo "$label: $rc"
  if [ $rc -eq 124 ] || [ $rc -eq 137 ]; then echo "ABORT: $label hung"; exit 99; fi
  return $rc
}
