"""Minimal .dat/.cod reader/writer for the TESTS (grammar: SURVEY.md appendix B,
reference datafile.c:112-148 header, 552-748 entry lines, 396-447 output).

Test helper only: the product's host layer has its own C reader."""
import numpy as np

TOPOL = {"data": 1, "lvq": 2, "hexa": 3, "rect": 4}
NEIGH = {"bubble": 1, "gaussian": 2}


class Entries:
    def __init__(self):
        self.dim = 0
        self.topol = "data"
        self.xdim = self.ydim = 0
        self.neigh = None
        self.points = None      # (n, dim) float32
        self.mask = None        # (n, dim) uint8 or None
        self.labels = []        # list of list of str

    @property
    def topol_id(self):
        return TOPOL[self.topol]

    @property
    def neigh_id(self):
        return NEIGH.get(self.neigh, 0)

    def first_label_ids(self, table=None):
        """first label of every entry as an int id (0 = no label), reference labels.h:45"""
        table = {} if table is None else table
        out = np.zeros(len(self.labels), np.int32)
        for i, l in enumerate(self.labels):
            if l:
                out[i] = table.setdefault(l[0], len(table) + 1)
        return out, table


def parse(text, mask_str="x"):
    e = Entries()
    rows, masks, labels = [], [], []
    header_done = False
    for line in text.splitlines():
        if line.startswith("#"):
            continue
        tok = line.replace("\t", " ").replace("\r", " ").split()
        if not tok:
            continue
        if not header_done:
            e.dim = int(tok[0])
            if len(tok) > 1:
                e.topol = tok[1]
            if len(tok) > 4:
                e.xdim, e.ydim, e.neigh = int(tok[2]), int(tok[3]), tok[4]
            header_done = True
            continue
        v = np.zeros(e.dim, np.float32)
        m = np.zeros(e.dim, np.uint8)
        for i in range(e.dim):
            if tok[i] == mask_str:
                m[i] = 1
            else:
                v[i] = np.float32(float(tok[i]))
        rows.append(v)
        masks.append(m)
        labels.append([t for t in tok[e.dim:] if not t.startswith("weight=") and not t.startswith("fixed=")])
    e.points = np.stack(rows) if rows else np.zeros((0, e.dim), np.float32)
    mk = np.stack(masks) if masks else np.zeros((0, e.dim), np.uint8)
    e.mask = mk if mk.any() else None
    e.labels = labels
    return e


def load(path):
    with open(path) as f:
        return parse(f.read())


def fmt_g(v):
    """C printf("%g") of a float promoted to double."""
    return "%g" % float(np.float32(v))


def format_entries(e, points=None, labels=None):
    """text exactly as write_header + write_entry produce it (datafile.c:396-447)"""
    pts = e.points if points is None else points
    labels = e.labels if labels is None else labels
    hdr = "%d" % e.dim
    if e.topol != "data":
        hdr += " %s" % e.topol
        if e.topol in ("hexa", "rect"):
            hdr += " %d %d %s" % (e.xdim, e.ydim, e.neigh)
    out = [hdr + "\n"]
    for i in range(pts.shape[0]):
        s = "".join(("x " if (e.mask is not None and e.mask[i, d]) else fmt_g(pts[i, d]) + " ")
                    for d in range(e.dim))
        s += "".join(l + " " for l in labels[i])
        out.append(s + "\n")
    return "".join(out)
