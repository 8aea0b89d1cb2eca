"""ILDG / LIME container on the host (src/output/ildg_format.jl in the reference: LIME_header, ILDG(filename), save_binarydata).

Only the container lives here: the binary payload goes to and from the GPU unchanged through
GaugeConfiguration.upload_ildg / to_ildg (gfb_gauge_upload_ildg / gfb_gauge_download_ildg), which do the byte swap, the precision
conversion and the layout transpose on the device.

LIME record = 144-byte header (magic 0x456789ab, version 1, flag bits MB = 0x8000 / ME = 0x4000, 64-bit big-endian data
length, 128-byte NUL-padded type string) + data padded to a multiple of 8 bytes.
"""
import re
import struct

LIME_MAGIC = 0x456789AB


def read_records(filename):
    """[(type, data bytes)] of a LIME file."""
    out = []
    with open(filename, "rb") as f:
        while True:
            head = f.read(144)
            if len(head) == 0:
                break
            if len(head) != 144:
                raise ValueError("truncated LIME header")
            magic, version, flags, length = struct.unpack(">IHHQ", head[:16])
            if magic != LIME_MAGIC:
                raise ValueError("bad LIME magic 0x%08x" % magic)
            rtype = head[16:144].split(b"\0", 1)[0].decode("ascii")
            data = f.read(length)
            if len(data) != length:
                raise ValueError("truncated LIME record %s" % rtype)
            f.read((-length) % 8)
            out.append((rtype, data))
    return out


def read_ildg(filename, lattice=None, precision=None):
    """(lattice (NX,NY,NZ,NT), precision, payload bytes, NC) of an ILDG file; NC follows from the payload size.
    Files without an ildg-format record (the reference's test/data fixtures are bare ildg-binary-data) need `lattice` and
    `precision` from the caller, as load_gaugefield!(U, i, ildg, L, NC) takes them in the reference."""
    fmt, payload = None, None
    for rtype, data in read_records(filename):
        if rtype == "ildg-format":
            fmt = data.decode("utf-8", "replace")
        elif rtype == "ildg-binary-data":
            payload = data
    if payload is None:
        raise ValueError("not an ILDG file: no ildg-binary-data record")

    def tag(name):
        m = re.search(r"<%s>\s*([^<\s]+)\s*</%s>" % (name, name), fmt)
        if not m:
            raise ValueError("ildg-format lacks <%s>" % name)
        return m.group(1)

    if fmt is not None:
        lattice = tuple(int(tag(n)) for n in ("lx", "ly", "lz", "lt"))
        precision = int(tag("precision"))
    elif lattice is None or precision is None:
        raise ValueError("no ildg-format record: pass lattice and precision")
    sites = lattice[0] * lattice[1] * lattice[2] * lattice[3]
    per_link = len(payload) // (sites * 4 * 2 * (precision // 8))
    nc = int(round(per_link ** 0.5))
    if nc * nc * sites * 4 * 2 * (precision // 8) != len(payload):
        raise ValueError("ildg-binary-data size does not match the lattice")
    return lattice, precision, payload, nc


def format_xml(lattice, precision, field="su3gauge"):
    return ('<?xml version="1.0" encoding="UTF-8"?>\n<ildgFormat xmlns="http://www.lqcd.org/ildg" '
            'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" xsi:schemaLocation="http://www.lqcd.org/ildg filefmt.xsd">\n'
            "  <version>1.0</version>\n  <field>%s</field>\n  <precision>%d</precision>\n"
            "  <lx>%d</lx>\n  <ly>%d</ly>\n  <lz>%d</lz>\n  <lt>%d</lt>\n</ildgFormat>\n"
            % (field, precision, lattice[0], lattice[1], lattice[2], lattice[3]))


def _record(rtype, data, begin, end):
    flags = (0x8000 if begin else 0) | (0x4000 if end else 0)
    head = struct.pack(">IHHQ", LIME_MAGIC, 1, flags, len(data)) + rtype.encode("ascii").ljust(128, b"\0")
    return head + data + b"\0" * ((-len(data)) % 8)


def write_ildg(filename, lattice, precision, payload):
    """One LIME message: ildg-format (XML) + ildg-binary-data."""
    with open(filename, "wb") as f:
        f.write(_record("ildg-format", format_xml(lattice, precision).encode("utf-8"), True, False))
        f.write(_record("ildg-binary-data", bytes(payload), False, True))


# ---- Bridge++ text format (src/output/bridge_format.jl:201-297: save_textdata / load_BridgeText!) --------------------------
# One decimal number per line, real then imaginary part, in the SAME element order as the ILDG payload
# ([t][z][y][x][mu][a][b]); so a text file is converted to/from the 64-bit ILDG payload and takes the same device path.
def read_bridge_text(filename, lattice, nc=3):
    """64-bit ILDG payload (bytes) of a Bridge++ text configuration; raises ValueError when the line count is wrong
    (the reference asserts 4*NX*NY*NZ*NT*NC*NC*2 lines)."""
    import numpy as np

    want = 4 * lattice[0] * lattice[1] * lattice[2] * lattice[3] * nc * nc * 2
    with open(filename) as f:
        vals = np.array([float(line.split()[0]) for line in f if line.strip()], dtype=np.float64)
    if vals.size != want:
        raise ValueError("Bridge text file has %d numbers; expected %d" % (vals.size, want))
    return vals.astype(">f8").tobytes()


def write_bridge_text(filename, payload):
    """Inverse of read_bridge_text for a 64-bit payload (repr round-trips a double exactly)."""
    import numpy as np

    vals = np.frombuffer(payload, dtype=">f8")
    with open(filename, "w") as f:
        f.write("\n".join(repr(float(v)) for v in vals))
        f.write("\n")
