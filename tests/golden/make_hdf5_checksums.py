#!/usr/bin/env python
"""Extract (metadata bytes, stored checksum) pairs from the reference's own HDF5 fixtures -> hdf5_checksums.json.
The checksums were written by libhdf5 when the reference authors created the files; they pin lookup3() in
fhmcanalysis_b200/io/hdf5_min.py without needing libhdf5 here.  Run in the build container (needs /root/reference)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
FILES = ["/root/reference/unittests/reference/test.nc", "/root/reference/unittests/reference/test2.nc",
         "/root/reference/example/ntot/square_well/T_0.90/composite.nc"]


def main():
    out = []
    for path in FILES:
        buf = open(path, "rb").read()
        name = os.path.relpath(path, "/root/reference")
        if buf[8] == 2:
            out.append({"what": name + " superblock v2", "bytes": buf[:44].hex(), "checksum": int.from_bytes(buf[44:48], "little")})
        pos, n_hdr = 0, 0
        while n_hdr < 3:
            pos = buf.find(b"OHDR\x02", pos)
            if pos < 0:
                break
            flags = buf[pos + 5]
            p = pos + 6 + (16 if flags & 0x20 else 0) + (4 if flags & 0x10 else 0)
            cs = 1 << (flags & 3)
            end = p + cs + int.from_bytes(buf[p:p + cs], "little")
            out.append({"what": "%s OHDR at %d" % (name, pos), "bytes": buf[pos:end].hex(),
                        "checksum": int.from_bytes(buf[end:end + 4], "little")})
            n_hdr += 1
            pos = end
    with open(os.path.join(HERE, "hdf5_checksums.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", len(out), "vectors")


if __name__ == "__main__":
    main()
