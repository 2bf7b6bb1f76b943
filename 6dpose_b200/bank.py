"""Template bank: storage, OpenCV-FileStorage YAML IO, and flat packing for the C-ABI.

Mirrors (reference: linemodLevelup/linemodLevelup.{h,cpp}, "LL.h"/"LL.cpp"):
  Feature / Template / TemplatePyramid / TemplatesMap     LL.h:23-45, 361-362
  Template::read/write, Feature::read/write               LL.cpp:194-232
  Detector::readClass / writeClass                        LL.cpp:2043-2122

A template pyramid is a list of L*M templates ordered [level*M + modality] (LL.cpp:1964).
Each template is (width, height, pyramid_level, features int32 [n,3] = x,y,label).

Packed bank files (no counterpart in the reference, whose only format is the
FileStorage YAML of LL.cpp:2093-2146): one little-endian binary file per class,

    "LMBANK1\0" | u32 n_templates, slots, pyramid_levels, name_len | u64 n_feats | name (padded to 8)
    int32 tmeta[n_templates][slots][4] = width, height, feat_begin (class-local), feat_count
    int16 x[n_feats] | int16 y[n_feats] | uint8 label[n_feats] (padded to 8) | u64 FNV-1a of all before

which loads with three frombuffer calls instead of a YAML parse (allScales, 2989 templates: see
tests/test_bank_packed.py) and stays packed in memory until somebody asks for Template objects.
"""
import os
import re
import struct

import numpy as np

MODALITY_NAMES = ["ColorGradient", "DepthNormal"]


class Template:
    __slots__ = ("width", "height", "pyramid_level", "features")

    def __init__(self, width=0, height=0, pyramid_level=0, features=None):
        self.width = int(width)
        self.height = int(height)
        self.pyramid_level = int(pyramid_level)
        self.features = np.zeros((0, 3), np.int32) if features is None else np.ascontiguousarray(features, np.int32).reshape(-1, 3)


PACKED_MAGIC = b"LMBANK1\0"
PACKED_SUFFIX = ".lmb"


def _fnv1a64(buf):
    """FNV-1a over 8-byte words (vectorised: xor-multiply chains do not vectorise, so the words are folded
    with a position-dependent multiplier instead -- still detects any flipped or truncated byte)."""
    a = np.frombuffer(buf, np.uint64)
    k = (np.arange(1, a.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
    with np.errstate(over="ignore"):
        return int(np.bitwise_xor.reduce(a * k) ^ np.uint64(0xCBF29CE484222325)) if a.size else 0xCBF29CE484222325


class PackedPyramids:
    """A class's template pyramids kept as flat arrays (what a packed bank file holds); behaves like the
    list of lists of Template the YAML reader builds, materialising Template objects on demand."""

    def __init__(self, tmeta, feats, levels):
        self.tmeta = np.ascontiguousarray(tmeta, np.int32)     # [n, slots, 4], feat_begin class-local
        self.feats = np.ascontiguousarray(feats, np.int32).reshape(-1, 3)
        self.levels = int(levels)
        self.modalities = self.tmeta.shape[1] // max(self.levels, 1)

    def __len__(self):
        return int(self.tmeta.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        tp = []
        for s_, (w, h, b, n) in enumerate(self.tmeta[i].tolist()):
            tp.append(Template(w, h, s_ // self.modalities, self.feats[b:b + n].copy()))
        return tp

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def append(self, tp):
        raise TypeError("packed class: convert with list(...) before adding templates")


class TemplateBank:
    """class_id -> list of template pyramids.  Iteration order for matching = std::map order
    (sorted class ids), as in LL.cpp:1756-1758."""

    def __init__(self):
        self.classes = {}

    def num_templates(self, class_id=None):
        if class_id is not None:
            return len(self.classes.get(class_id, []))
        return sum(len(v) for v in self.classes.values())

    def class_ids(self):
        return sorted(self.classes.keys())

    # ---- YAML (OpenCV FileStorage dialect, YAML 1.0) -------------------------------------
    _tok = re.compile(
        r"^\s*(?:-\s*)?(?:"
        r"(?P<key>class_id|pyramid_levels|template_id|width|height|pyramid_level|modalities)\s*:\s*(?P<val>.*?)\s*$"
        r"|\[\s*(?P<x>-?\d+)\s*,\s*(?P<y>-?\d+)\s*,\s*(?P<l>-?\d+)\s*\]\s*$)",
        re.M)

    def read_class(self, path, expected_levels, class_id_override=""):
        """Detector::readClass, LL.cpp:2043-2091.  Raises RuntimeError where the reference CV_Asserts."""
        import gzip
        opener = gzip.open if path.endswith(".gz") else open
        try:
            with opener(path, "rt") as fh:
                text = fh.read()
        except OSError as e:
            raise RuntimeError("cannot open template file %s: %s" % (path, e))
        class_id = None
        levels = None
        pyramids = []
        cur_tp = None
        cur_t = None
        feats = None
        for m in self._tok.finditer(text):
            if m.group("x") is not None:
                feats.append((int(m.group("x")), int(m.group("y")), int(m.group("l"))))
                continue
            key, val = m.group("key"), m.group("val")
            if key == "class_id":
                class_id = val.strip().strip('"')
            elif key == "modalities":
                names = [s.strip() for s in val.strip("[] ").split(",") if s.strip()]
                if names != MODALITY_NAMES:  # LL.cpp:2047-2051
                    raise RuntimeError("modalities mismatch: %r" % (names,))
            elif key == "pyramid_levels":
                levels = int(val)
            elif key == "template_id":
                if int(val) != len(pyramids):  # LL.cpp:2077
                    raise RuntimeError("template_id == expected_id")
                cur_tp = []
                pyramids.append(cur_tp)
            elif key == "width":
                cur_t = Template(width=int(val))
                feats = []
                cur_t.features = feats
                cur_tp.append(cur_t)
            elif key == "height":
                cur_t.height = int(val)
            elif key == "pyramid_level":
                cur_t.pyramid_level = int(val)
        if levels != expected_levels:  # LL.cpp:2052
            raise RuntimeError("pyramid_levels mismatch: file %r detector %r" % (levels, expected_levels))
        if class_id_override:
            class_id = class_id_override
        elif class_id in self.classes:  # LL.cpp:2059
            raise RuntimeError("class %s already loaded" % class_id)
        for tp in pyramids:
            for t in tp:
                t.features = np.asarray(t.features, np.int32).reshape(-1, 3)
        self.classes[class_id] = pyramids
        return class_id

    def write_class(self, class_id, path, levels):
        """Detector::writeClass, LL.cpp:2093-2122 (same node layout as the reference's files)."""
        import gzip
        tps = self.classes[class_id]
        out = ["%YAML:1.0", "---", 'class_id: "%s"' % class_id,
               "modalities: [ %s ]" % ", ".join(MODALITY_NAMES), "pyramid_levels: %d" % levels,
               "template_pyramids:"]
        for i, tp in enumerate(tps):
            out.append("   -")
            out.append("      template_id: %d" % i)
            out.append("      templates:")
            for t in tp:
                out.append("         -")
                out.append("            width: %d" % t.width)
                out.append("            height: %d" % t.height)
                out.append("            pyramid_level: %d" % t.pyramid_level)
                out.append("            features:")
                for x, y, l in t.features.tolist():
                    out.append("               - [ %d, %d, %d ]" % (x, y, l))
        data = "\n".join(out) + "\n"
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "wt") as fh:
            fh.write(data)

    # ---- packed binary files (section 8f-4) -----------------------------------------------
    def write_packed(self, class_id, path, levels):
        one = self.pack([class_id], None)
        tmeta, feats = one["tmeta"], one["feats"]
        slots = tmeta.shape[1] if tmeta.size else levels * len(MODALITY_NAMES)
        if feats.size and (np.abs(feats[:, :2]).max() > 32767 or feats[:, 2].min() < 0 or feats[:, 2].max() > 255):
            raise RuntimeError("feature coordinates / labels outside the packed format's int16 / uint8 range")
        for tp in (self.classes[class_id] if not isinstance(self.classes[class_id], PackedPyramids) else []):
            for s_, t in enumerate(tp):
                if t.pyramid_level != s_ // len(MODALITY_NAMES):
                    raise RuntimeError("pyramid_level does not follow the slot order (LL.cpp:1964)")
        name = class_id.encode("utf-8")
        pad = lambda b: b + b"\0" * (-len(b) % 8)
        body = b"".join([
            PACKED_MAGIC, struct.pack("<IIIIQ", tmeta.shape[0], slots, levels, len(name), feats.shape[0]), pad(name),
            tmeta.astype("<i4").tobytes(), pad(feats[:, 0].astype("<i2").tobytes()), pad(feats[:, 1].astype("<i2").tobytes()),
            pad(feats[:, 2].astype("u1").tobytes())])
        tmp = path + ".tmp%d" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(body)
            fh.write(struct.pack("<Q", _fnv1a64(body)))
        os.replace(tmp, path)

    def read_packed(self, path, expected_levels, class_id_override=""):
        """Same contract as read_class (LL.cpp:2043-2091) on a packed file."""
        try:
            with open(path, "rb") as fh:
                blob = fh.read()
        except OSError as e:
            raise RuntimeError("cannot open template file %s: %s" % (path, e))
        if len(blob) < 40 or blob[:8] != PACKED_MAGIC or len(blob) % 8:
            raise RuntimeError("%s is not a packed template bank" % path)
        body, (digest,) = blob[:-8], struct.unpack("<Q", blob[-8:])
        if _fnv1a64(body) != digest:
            raise RuntimeError("%s: checksum mismatch (truncated or corrupted)" % path)
        n, slots, levels, name_len, nf = struct.unpack("<IIIIQ", body[8:32])
        if levels != expected_levels:  # LL.cpp:2052
            raise RuntimeError("pyramid_levels mismatch: file %r detector %r" % (levels, expected_levels))
        if slots != levels * len(MODALITY_NAMES):
            raise RuntimeError("modalities mismatch: %d slots for %d levels" % (slots, levels))
        r8 = lambda v: (v + 7) & ~7
        o = 32
        class_id = body[o:o + name_len].decode("utf-8")
        o += r8(name_len)
        need = o + n * slots * 16 + 2 * r8(2 * nf) + r8(nf)
        if need != len(body):
            raise RuntimeError("%s: size does not match its header" % path)
        tmeta = np.frombuffer(body, "<i4", n * slots * 4, o).reshape(n, slots, 4)
        o += n * slots * 16
        fx = np.frombuffer(body, "<i2", nf, o)
        o += r8(2 * nf)
        fy = np.frombuffer(body, "<i2", nf, o)
        o += r8(2 * nf)
        fl = np.frombuffer(body, "u1", nf, o)
        if n and (tmeta[:, :, 3].min() < 0 or tmeta[:, :, 2].min() < 0 or int((tmeta[:, :, 2] + tmeta[:, :, 3]).max()) > nf):
            raise RuntimeError("%s: feature ranges outside the file" % path)
        if class_id_override:
            class_id = class_id_override
        elif class_id in self.classes:  # LL.cpp:2059
            raise RuntimeError("class %s already loaded" % class_id)
        feats = np.empty((nf, 3), np.int32)
        feats[:, 0], feats[:, 1], feats[:, 2] = fx, fy, fl
        self.classes[class_id] = PackedPyramids(tmeta, feats, levels)
        return class_id

    def read_any(self, path, expected_levels, cache=None):
        """Packed file if `path` names one; else the YAML, through a packed sibling (`path + ".lmb"`, rebuilt
        when older than the YAML) when `cache` (default: env LINEMOD_B200_BANK_CACHE=1) asks for it."""
        if path.endswith(PACKED_SUFFIX):
            return self.read_packed(path, expected_levels)
        if cache is None:
            cache = os.environ.get("LINEMOD_B200_BANK_CACHE", "0") == "1"
        side = path + PACKED_SUFFIX
        if cache and os.path.exists(side) and os.path.exists(path) and os.path.getmtime(side) >= os.path.getmtime(path):
            try:
                return self.read_packed(side, expected_levels)
            except RuntimeError as e:
                if "already loaded" in str(e) or "mismatch: file" in str(e):
                    raise
        cid = self.read_class(path, expected_levels)
        if cache:
            try:
                self.write_packed(cid, side, expected_levels)
            except OSError:
                pass  # read-only location: keep parsing the YAML
        return cid

    # ---- flat packing ------------------------------------------------------------------
    def pack(self, class_ids, slots):
        """Flatten the given classes (in the given order) for the C-ABI / oracle.

        Returns dict(class_begin int32 [C+1], tmeta int32 [G, slots, 4] = width,height,feat_begin,feat_count,
        feats int32 [F,3]).  slots = pyramid_levels * modalities.
        """
        class_begin = [0]
        metas = []
        feats = []
        nfe = 0
        for cid in class_ids:
            tps = self.classes[cid]
            if isinstance(tps, PackedPyramids):
                if slots is not None and len(tps) and tps.tmeta.shape[1] != slots:
                    raise RuntimeError("template pyramid of class %s has %d templates, expected %d" % (cid, tps.tmeta.shape[1], slots))
                tm = tps.tmeta.copy()
                tm[:, :, 2] += nfe
                metas.append(tm)
                feats.append(tps.feats)
                nfe += int(tps.feats.shape[0])
                class_begin.append(class_begin[-1] + len(tps))
                continue
            rows0 = len(metas)
            for tp in tps:
                if slots is None:
                    slots = len(tp)
                if len(tp) != slots:
                    raise RuntimeError("template pyramid of class %s has %d templates, expected %d" % (cid, len(tp), slots))
                row = []
                for t in tp:
                    n = int(t.features.shape[0])
                    row.append((t.width, t.height, nfe, n))
                    feats.append(t.features)
                    nfe += n
                metas.append(np.asarray(row, np.int32).reshape(1, slots, 4))
            class_begin.append(class_begin[-1] + len(metas) - rows0)
        if slots is None:
            slots = metas[0].shape[1] if metas else 0
        tmeta = np.concatenate(metas, 0).astype(np.int32) if metas else np.zeros((0, slots, 4), np.int32)
        allf = np.concatenate(feats, 0).astype(np.int32) if feats else np.zeros((0, 3), np.int32)
        return dict(class_begin=np.asarray(class_begin, np.int32), tmeta=np.ascontiguousarray(tmeta),
                    feats=np.ascontiguousarray(allf))
