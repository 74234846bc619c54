#!/bin/sh
# Regenerates the known-answer values quoted in tests/golden/README.md with the unmodified reference binary
# (oracle/_ref/bzip3_ref, built from /root/reference by `make -C oracle ref`).
set -e
here=$(cd "$(dirname "$0")" && pwd)
ref="$here/../../oracle/_ref/bzip3_ref"
"$ref" -e -b 8 < "$here/shakespeare.txt" > /tmp/kat_b8.bz3
wc -c /tmp/kat_b8.bz3
sha256sum /tmp/kat_b8.bz3
python3 - <<'PY'
import struct
b = open('/tmp/kat_b8.bz3', 'rb').read()
cs, osz = struct.unpack('<ii', b[9:17])
crc, idx, model, lzp = struct.unpack('<IiBi', b[17:30])
print('csize', cs, 'orig', osz, 'crc', hex(crc), 'bwt_idx', idx, 'model', model, 'lzp_size', lzp)
PY
"$ref" -d < "$here/shakespeare.txt.bz3" | cmp - "$here/shakespeare.txt" && echo "golden decode ok"
