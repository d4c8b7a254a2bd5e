#!/usr/bin/env python3
"""Host utilities of the reference that sit next to the hot path, over the
library's own host layer:

    lz4ada_tools.py xxhash32 < file      counterpart of tool_xxhash32ada/xxhash32ada.adb:15-26
    lz4ada_tools.py hdrinfo  < file.lz4  counterpart of tool_lz4hdrinfo/lz4hdrinfo.adb:70-145

Neither needs a GPU: the frame-header parser and the public XXHash32 API are host code."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bo_lz4_ada_b200 as lz  # noqa: E402


def xxhash32():
    h = lz.XXHash32.Init()
    while True:
        chunk = sys.stdin.buffer.read(1 << 16)
        if not chunk:
            break
        h.Update(chunk)
    print(lz.To_Hex(h.Final()))


def hdrinfo():
    data = sys.stdin.buffer.read(4096)
    try:
        dec, consumed = lz.Init_With_Header(data, "Use_First")
    except lz.LZ4AdaError as e:
        print(e.information)
        return 1
    magic = int.from_bytes(data[:4], "little")
    print("magic            0x%08x" % magic)
    print("header bytes     %d" % consumed)
    print("min buffer size  %d  (block max + 64 KiB + 8)" % dec.Min_Buffer_Size)
    if magic == 0x184D2204:
        flg, bd = data[4], data[5]
        print("block max        %d" % (dec.Min_Buffer_Size - 65536 - 8))
        print("block independ.  %s" % bool(flg & 0x20))
        print("block checksum   %s" % bool(flg & 0x10))
        print("content size     %s" % (int.from_bytes(data[6:14], "little") if flg & 8 else "absent"))
        print("content checksum %s" % bool(flg & 4))
        print("dictionary id    %s" % bool(flg & 1))
        print("header checksum  0x%02x" % data[consumed - 1])
    elif magic == 0x184C2102:
        print("legacy frame (8 MiB blocks, no checksums)")
    else:
        print("skippable frame")
    print("end of frame     %s" % dec.Is_End_Of_Frame())
    return 0


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else ""
    sys.exit({"xxhash32": xxhash32, "hdrinfo": hdrinfo}.get(cmd, lambda: print(__doc__) or 2)() or 0)
