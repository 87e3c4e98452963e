"""ctypes mirror of ``include/jpp_b200.h`` plus numpy-backed holders for the
model description, staged lattice batches and results.

The structures here are byte-for-byte the C ABI; nothing in this module
computes scores.  ``StagedBatch`` / ``ModelSpec`` can be built from the JPPB
files that the reference-side stager writes (``oracle/ref_harness.cc``) or
from arrays produced by any other host-side lattice builder.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from dataclasses import dataclass, field

import numpy as np

ABI_VERSION = 3

# jpp_b200_status
OK, INVALID_PARAMETER, INVALID_STATE, NOT_IMPLEMENTED, CUDA_ERROR = range(5)

# jpp_b200_float_order
ORDER_STATIC_CODEGEN, ORDER_DYNAMIC = 0, 1

# jpp_b200_sentence_status bits
SENT_BAD_OFFSETS, SENT_BAD_SHAPE, SENT_BAD_ENDS, SENT_BAD_SPAN, SENT_BAD_ENTRY, SENT_BAD_RNN_ID = 1, 2, 4, 8, 16, 32

# jpp_b200_prim_kind
PRIM_COPY, PRIM_SHIFT_MASK, PRIM_PROVIDED, PRIM_BYTE_LENGTH, PRIM_CODEPOINT_SIZE = 1, 2, 3, 4, 5
PRIM_SURFACE_CODEPOINT_SIZE, PRIM_CODEPOINT_TYPE, PRIM_CODEPOINT = 6, 7, 8


class Primitive(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("c", C.c_int32)]


class Computation(C.Structure):
    _fields_ = [("primitive", C.c_int32), ("true_off", C.c_uint32), ("true_len", C.c_uint32),
                ("false_off", C.c_uint32), ("false_len", C.c_uint32)]


class Pattern(C.Structure):
    _fields_ = [("index", C.c_uint32), ("ref_off", C.c_uint32), ("ref_len", C.c_uint32)]


class Ngram(C.Structure):
    _fields_ = [("order", C.c_uint32), ("index", C.c_uint32), ("refs", C.c_uint32 * 3)]


class RnnModel(C.Structure):
    _fields_ = [("layer_size", C.c_uint32), ("maxent_order", C.c_uint32), ("maxent_size", C.c_uint64),
                ("vocab_size", C.c_uint64), ("nce_constant", C.c_float), ("unk_id", C.c_int32),
                ("unk_constant", C.c_float), ("unk_length_penalty", C.c_float),
                ("matrix", C.c_void_p), ("embeddings", C.c_void_p), ("nce_embeddings", C.c_void_p),
                ("maxent", C.c_void_p)]


class Model(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("weights", C.c_void_p), ("log2_weights", C.c_uint32),
                ("weight_seed", C.c_uint64), ("weight_scale", C.c_float),
                ("entry_size", C.c_uint32), ("num_placeholders", C.c_uint32),
                ("num_primitives", C.c_uint32), ("primitives", C.POINTER(Primitive)),
                ("num_computations", C.c_uint32), ("computations", C.POINTER(Computation)),
                ("branch_pool_size", C.c_uint32), ("branch_pool", C.c_void_p),
                ("num_patterns", C.c_uint32), ("num_stored_patterns", C.c_uint32),
                ("patterns", C.POINTER(Pattern)),
                ("pattern_ref_pool_size", C.c_uint32), ("pattern_ref_pool", C.c_void_p),
                ("num_ngrams", C.c_uint32), ("ngrams", C.POINTER(Ngram)),
                ("num_scorers", C.c_uint32), ("score_weights", C.c_float * 2),
                ("rnn", C.POINTER(RnnModel)), ("float_order", C.c_uint32)]


class Config(C.Structure):
    _fields_ = [("beam_size", C.c_int32), ("global_beam_size", C.c_int32),
                ("right_gbeam_check", C.c_int32), ("right_gbeam_size", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("num_sentences", C.c_uint32),
                ("sent_boundary_off", C.c_void_p), ("sent_codepoint_off", C.c_void_p),
                ("boundary_node_off", C.c_void_p), ("boundary_end_off", C.c_void_p),
                ("node_info", C.c_void_p), ("node_entry", C.c_void_p), ("node_placeholders", C.c_void_p),
                ("node_rnn_id", C.c_void_p), ("ends", C.c_void_p), ("codepoints", C.c_void_p),
                ("node_entry_row", C.c_void_p), ("entry_rows", C.c_void_p), ("num_entry_rows", C.c_uint32),
                ("row_placeholders", C.c_void_p), ("row_rnn_id", C.c_void_p), ("sent_extra_row_off", C.c_void_p)]


class EntryTableStruct(C.Structure):
    _fields_ = [("num_rows", C.c_uint32), ("entry_ptr", C.c_void_p), ("rows", C.c_void_p), ("rnn_id", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("node_beam", C.c_void_p), ("node_beam_scores", C.c_void_p),
                ("node_patterns", C.c_void_p), ("node_t0_score", C.c_void_p),
                ("best_path_len", C.c_void_p), ("best_path_nodes", C.c_void_p), ("sentence_status", C.c_void_p)]


class Counters(C.Structure):
    _fields_ = [("sentences", C.c_uint64), ("boundaries", C.c_uint64), ("nodes", C.c_uint64),
                ("gathers", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("rnn_nodes", C.c_uint64), ("rnn_scores", C.c_uint64), ("walk_gathers", C.c_uint64),
                ("walk_bytes", C.c_uint64), ("launches", C.c_uint64), ("kernel_ms", C.c_float * 8),
                ("walk_gathers_issued", C.c_uint64), ("failed_sentences", C.c_uint64)]


NODE_INFO_DT = np.dtype([("entry_ptr", "<i4"), ("start", "<u2"), ("end", "<u2")])
NODE_PTR_DT = np.dtype([("boundary", "<u2"), ("position", "<u2")])
CODEPOINT_DT = np.dtype([("codepoint", "<u4"), ("char_class", "<u4")])
BEAM_EL_DT = np.dtype([("left", "<u2"), ("prev_beam", "<u2"), ("total_score", "<f4")])


def _ptr(a: np.ndarray | None) -> int | None:
    return None if a is None else a.ctypes.data


@dataclass
class RnnSpec:
    layer_size: int
    maxent_order: int
    maxent_size: int
    vocab_size: int
    nce_constant: float
    unk_id: int
    unk_constant: float
    unk_length_penalty: float
    matrix: np.ndarray
    embeddings: np.ndarray
    nce_embeddings: np.ndarray
    maxent: np.ndarray

    @staticmethod
    def from_files(prefix: str, icfg: np.ndarray, fcfg: np.ndarray) -> "RnnSpec":
        """Read the four float blocks of a ``.nnet`` v6 file
        (reference src/rnn/mikolov_rnn.cc:29-76 header, :196-209 blocks)."""
        h, order, msize, vocab, unk_id = (int(x) for x in icfg[:5])
        raw = np.fromfile(prefix + ".nnet", dtype=np.uint8)
        off = 8 + 8 + 4 + 1 + 4 + 1 + 64 + 4 + 4
        f = raw[off:].view(np.float32) if (len(raw) - off) % 4 == 0 else np.frombuffer(raw[off:].tobytes(), np.float32)
        emb = f[: vocab * h]
        nce = f[vocab * h: 2 * vocab * h]
        mat = f[2 * vocab * h: 2 * vocab * h + h * h]
        me = f[2 * vocab * h + h * h: 2 * vocab * h + h * h + msize]
        assert me.size == msize, "truncated .nnet"
        return RnnSpec(h, order, msize, vocab, float(fcfg[0]), unk_id, float(fcfg[1]), float(fcfg[2]),
                       np.ascontiguousarray(mat), np.ascontiguousarray(emb), np.ascontiguousarray(nce),
                       np.ascontiguousarray(me))


@dataclass
class ModelSpec:
    """Compiled feature program + scorer constants (jpp_b200_model)."""
    log2_weights: int
    weight_seed: int
    weight_scale: float
    entry_size: int
    num_placeholders: int
    prims: np.ndarray          # [n][4] i32
    comps: np.ndarray          # [n][5] i32
    branch_pool: np.ndarray    # i32
    patterns: np.ndarray       # [n][3] i32
    pattern_ref_pool: np.ndarray
    ngrams: np.ndarray         # [n][5] i32: order, index, refs[3]
    num_stored_patterns: int
    score_weights: tuple[float, float] = (1.0, 0.0)
    num_scorers: int = 1
    weights: np.ndarray | None = None   # explicit table, else synthesised from the seed
    rnn: RnnSpec | None = None
    float_order: int = ORDER_STATIC_CODEGEN
    _keep: list = field(default_factory=list, repr=False)

    @staticmethod
    def from_jppb(d: dict[str, np.ndarray], rnn_prefix: str | None = None) -> "ModelSpec":
        cfg, fcfg, cnt = d["cfg"], d["fcfg"], d["spec_counts"]
        ms = ModelSpec(
            log2_weights=int(cfg[5]), weight_seed=int(cfg[6]), weight_scale=float(fcfg[0]),
            entry_size=int(cnt[5]), num_placeholders=int(cnt[6]),
            prims=np.ascontiguousarray(d["spec_prims"], np.int32),
            comps=np.ascontiguousarray(d["spec_comps"], np.int32),
            branch_pool=np.ascontiguousarray(d["spec_branch_pool"], np.int32),
            patterns=np.ascontiguousarray(d["spec_patterns"], np.int32),
            pattern_ref_pool=np.ascontiguousarray(d["spec_pattern_ref_pool"], np.int32),
            ngrams=np.ascontiguousarray(d["spec_ngrams"], np.int32),
            num_stored_patterns=int(cnt[3]),
            score_weights=(float(fcfg[1]), float(fcfg[2])), num_scorers=int(cfg[4]),
            # staged from the reference's dynamic feature path (harness dynamic=1): its float order
            float_order=ORDER_DYNAMIC if len(cfg) > 8 and int(cfg[8]) else ORDER_STATIC_CODEGEN)
        if "weights" in d:
            ms.weights = np.ascontiguousarray(d["weights"], np.float32)
        if ms.num_scorers == 2:
            if rnn_prefix is None:
                raise ValueError("model has an RNN scorer: pass rnn_prefix")
            ms.rnn = RnnSpec.from_files(rnn_prefix, d["rnn_icfg"], d["rnn_fcfg"])
        return ms

    def as_struct(self) -> Model:
        m = Model()
        m.abi_version = ABI_VERSION
        m.weights = _ptr(self.weights)
        m.log2_weights = self.log2_weights
        m.weight_seed = self.weight_seed
        m.weight_scale = self.weight_scale
        m.entry_size = self.entry_size
        m.num_placeholders = self.num_placeholders
        prims = np.ascontiguousarray(self.prims, np.int32)
        comps = np.ascontiguousarray(self.comps, np.int32).astype(np.int32)
        pats = np.ascontiguousarray(self.patterns, np.int32)
        ngr = np.ascontiguousarray(self.ngrams, np.int32)
        self._keep = [prims, comps, pats, ngr]
        m.num_primitives = len(prims)
        m.primitives = C.cast(prims.ctypes.data, C.POINTER(Primitive))
        m.num_computations = len(comps)
        m.computations = C.cast(comps.ctypes.data, C.POINTER(Computation))
        m.branch_pool_size = len(self.branch_pool)
        m.branch_pool = _ptr(self.branch_pool)
        m.num_patterns = len(pats)
        m.num_stored_patterns = self.num_stored_patterns
        m.patterns = C.cast(pats.ctypes.data, C.POINTER(Pattern))
        m.pattern_ref_pool_size = len(self.pattern_ref_pool)
        m.pattern_ref_pool = _ptr(self.pattern_ref_pool)
        m.num_ngrams = len(ngr)
        m.ngrams = C.cast(ngr.ctypes.data, C.POINTER(Ngram))
        m.num_scorers = self.num_scorers
        m.score_weights[0] = self.score_weights[0]
        m.score_weights[1] = self.score_weights[1]
        m.float_order = self.float_order
        if self.rnn is not None:
            r = RnnModel()
            r.layer_size, r.maxent_order = self.rnn.layer_size, self.rnn.maxent_order
            r.maxent_size, r.vocab_size = self.rnn.maxent_size, self.rnn.vocab_size
            r.nce_constant, r.unk_id = self.rnn.nce_constant, self.rnn.unk_id
            r.unk_constant, r.unk_length_penalty = self.rnn.unk_constant, self.rnn.unk_length_penalty
            r.matrix, r.embeddings = _ptr(self.rnn.matrix), _ptr(self.rnn.embeddings)
            r.nce_embeddings, r.maxent = _ptr(self.rnn.nce_embeddings), _ptr(self.rnn.maxent)
            self._keep.append(r)
            m.rnn = C.pointer(r)
        return m


@dataclass
class ScoringConfig:
    beam_size: int = 5
    global_beam_size: int = 6
    right_gbeam_check: int = 1
    right_gbeam_size: int = 5

    @staticmethod
    def from_jppb(d: dict[str, np.ndarray]) -> "ScoringConfig":
        c = d["cfg"]
        return ScoringConfig(int(c[0]), int(c[1]), int(c[2]), int(c[3]))

    def as_struct(self) -> Config:
        return Config(self.beam_size, self.global_beam_size, self.right_gbeam_check, self.right_gbeam_size)


@dataclass
class StagedBatch:
    """N built lattices in CSR form (jpp_b200_batch)."""
    sent_boundary_off: np.ndarray
    sent_codepoint_off: np.ndarray
    boundary_node_off: np.ndarray
    boundary_end_off: np.ndarray
    node_info: np.ndarray        # NODE_INFO_DT
    node_entry: np.ndarray | None   # [nodes][entry_size] i32, or None in the compact form
    node_placeholders: np.ndarray | None
    node_rnn_id: np.ndarray | None
    ends: np.ndarray             # NODE_PTR_DT
    codepoints: np.ndarray       # CODEPOINT_DT
    node_entry_row: np.ndarray | None = None   # compact form: [nodes] u32 index into entry_rows
    entry_rows: np.ndarray | None = None       # [rows][entry_size] i32
    row_placeholders: np.ndarray | None = None # [rows][placeholders] i32 (then node_placeholders is None)
    row_rnn_id: np.ndarray | None = None       # [rows] i32 (then node_rnn_id is None)
    sent_extra_row_off: np.ndarray | None = None   # resident form: [sentences + 1] u32, rows = the sentences' extra nodes

    @property
    def num_sentences(self) -> int:
        return len(self.sent_boundary_off) - 1

    @property
    def num_nodes(self) -> int:
        return int(self.boundary_node_off[-1])

    @property
    def num_boundaries(self) -> int:
        return len(self.boundary_node_off) - 1

    @staticmethod
    def from_jppb(d: dict[str, np.ndarray]) -> "StagedBatch":
        n = len(d["node_entry_ptr"])
        ni = np.empty(n, NODE_INFO_DT)
        ni["entry_ptr"] = d["node_entry_ptr"]
        ni["start"] = d["node_start_end"][:, 0]
        ni["end"] = d["node_start_end"][:, 1]
        ends = np.ascontiguousarray(d["ends"], np.uint16).view(NODE_PTR_DT).reshape(-1)
        cps = np.ascontiguousarray(d["codepoints"], np.uint32).view(CODEPOINT_DT).reshape(-1)
        return StagedBatch(
            np.ascontiguousarray(d["sent_boundary_off"], np.uint32),
            np.ascontiguousarray(d["sent_codepoint_off"], np.uint32),
            np.ascontiguousarray(d["boundary_node_off"], np.uint32),
            np.ascontiguousarray(d["boundary_end_off"], np.uint32),
            ni, np.ascontiguousarray(d["node_entry"], np.int32),
            np.ascontiguousarray(d["node_placeholders"], np.int32) if "node_placeholders" in d else None,
            np.ascontiguousarray(d["node_rnn_id"], np.int32) if "node_rnn_id" in d else None,
            ends, cps)

    def as_struct(self) -> Batch:
        b = Batch()
        b.num_sentences = self.num_sentences
        b.sent_boundary_off = _ptr(self.sent_boundary_off)
        b.sent_codepoint_off = _ptr(self.sent_codepoint_off)
        b.boundary_node_off = _ptr(self.boundary_node_off)
        b.boundary_end_off = _ptr(self.boundary_end_off)
        b.node_info = _ptr(self.node_info)
        b.node_entry = _ptr(self.node_entry)
        b.node_entry_row = _ptr(self.node_entry_row)
        b.entry_rows = _ptr(self.entry_rows)
        b.num_entry_rows = 0 if self.entry_rows is None else len(self.entry_rows)
        if self.sent_extra_row_off is not None:
            b.num_entry_rows = int(self.sent_extra_row_off[-1])
        b.row_placeholders = _ptr(self.row_placeholders)
        b.row_rnn_id = _ptr(self.row_rnn_id)
        b.sent_extra_row_off = _ptr(self.sent_extra_row_off)
        b.node_placeholders = _ptr(self.node_placeholders)
        b.node_rnn_id = _ptr(self.node_rnn_id)
        b.ends = _ptr(self.ends)
        b.codepoints = _ptr(self.codepoints)
        return b

    @staticmethod
    def concat(parts: list["StagedBatch"], pin=None) -> "StagedBatch":
        """Concatenate batches (offsets rebased). `pin(array)` may copy into pinned memory."""
        pin = pin or (lambda a: a)

        def offs(get, base_of):
            out = [np.zeros(1, np.uint32)]
            base = 0
            for p in parts:
                o = get(p)
                out.append((o[1:].astype(np.uint64) + base).astype(np.uint32))
                base += int(o[-1])
            return np.concatenate(out)

        def cat(get):
            arrs = [get(p) for p in parts]
            return None if arrs[0] is None else pin(np.concatenate(arrs))

        rows = row_idx = row_ph = row_rnn = None
        if parts[0].node_entry is None:   # compact parts: stack the row tables, rebase the indices
            bases = np.cumsum([0] + [len(p.entry_rows) for p in parts[:-1]])
            rows = pin(np.concatenate([p.entry_rows for p in parts]))
            row_idx = pin(np.concatenate([(p.node_entry_row.astype(np.uint64) + int(b)).astype(np.uint32)
                                          for p, b in zip(parts, bases)]))
            if parts[0].row_placeholders is not None:
                row_ph = pin(np.concatenate([p.row_placeholders for p in parts]))
            if parts[0].row_rnn_id is not None:
                row_rnn = pin(np.concatenate([p.row_rnn_id for p in parts]))
        return StagedBatch(
            pin(offs(lambda p: p.sent_boundary_off, None)), pin(offs(lambda p: p.sent_codepoint_off, None)),
            pin(offs(lambda p: p.boundary_node_off, None)), pin(offs(lambda p: p.boundary_end_off, None)),
            cat(lambda p: p.node_info), cat(lambda p: p.node_entry), cat(lambda p: p.node_placeholders),
            cat(lambda p: p.node_rnn_id), cat(lambda p: p.ends), cat(lambda p: p.codepoints), row_idx, rows, row_ph, row_rnn)

    def select(self, lo: int, hi: int) -> "StagedBatch":
        """Sentences [lo, hi) as an independent batch (offsets rebased)."""
        b0, b1 = int(self.sent_boundary_off[lo]), int(self.sent_boundary_off[hi])
        n0, n1 = int(self.boundary_node_off[b0]), int(self.boundary_node_off[b1])
        e0, e1 = int(self.boundary_end_off[b0]), int(self.boundary_end_off[b1])
        c0, c1 = int(self.sent_codepoint_off[lo]), int(self.sent_codepoint_off[hi])
        return StagedBatch(
            (self.sent_boundary_off[lo:hi + 1] - b0).astype(np.uint32),
            (self.sent_codepoint_off[lo:hi + 1] - c0).astype(np.uint32),
            (self.boundary_node_off[b0:b1 + 1] - n0).astype(np.uint32),
            (self.boundary_end_off[b0:b1 + 1] - e0).astype(np.uint32),
            np.ascontiguousarray(self.node_info[n0:n1]),
            None if self.node_entry is None else np.ascontiguousarray(self.node_entry[n0:n1]),
            None if self.node_placeholders is None else np.ascontiguousarray(self.node_placeholders[n0:n1]),
            None if self.node_rnn_id is None else np.ascontiguousarray(self.node_rnn_id[n0:n1]),
            np.ascontiguousarray(self.ends[e0:e1]), np.ascontiguousarray(self.codepoints[c0:c1]),
            None if self.node_entry_row is None else np.ascontiguousarray(self.node_entry_row[n0:n1]),
            self.entry_rows, self.row_placeholders, self.row_rnn_id)

    def compact_entries(self, pin=None, fold: bool = True) -> "StagedBatch":
        """The same batch with one row per distinct entry (jpp_b200_batch.node_entry_row / entry_rows):
        what a stager that keys rows by EntryPtr sends.  With `fold` the placeholders and the RNN id --
        functions of the entry too -- ride in the row (row_placeholders / row_rnn_id) and the per-node
        arrays are dropped.  Rows are matched by content here, through a 64-bit hash whose collisions are
        checked.  A batch that is already compact (e.g. a concatenation of compact parts) gets its row
        table de-duplicated once more."""
        pin = pin or (lambda a: a)
        was_inline = self.node_entry is not None
        cols = [np.ascontiguousarray(self.node_entry if was_inline else self.entry_rows, np.int32)]
        es = cols[0].shape[1]
        ph = self.node_placeholders if was_inline or self.row_placeholders is None else self.row_placeholders
        rid = self.node_rnn_id if was_inline or self.row_rnn_id is None else self.row_rnn_id
        fold_ph = ph is not None and (self.row_placeholders is not None or (fold and was_inline))
        fold_rnn = rid is not None and (self.row_rnn_id is not None or (fold and was_inline))
        nph = 0
        if fold_ph:
            nph = ph.shape[1]
            cols.append(np.ascontiguousarray(ph, np.int32).reshape(len(ph), nph))
        if fold_rnn:
            cols.append(np.ascontiguousarray(rid, np.int32).reshape(-1, 1))
        rows = np.ascontiguousarray(np.concatenate(cols, axis=1)) if len(cols) > 1 else cols[0]
        n, width = rows.shape
        h = np.full(n, 0x9E3779B97F4A7C15, np.uint64)
        with np.errstate(over="ignore"):
            for c in range(width):
                h ^= (rows[:, c].astype(np.uint32).astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
                      + (h << np.uint64(6)) + (h >> np.uint64(2)))
                h *= np.uint64(0xFF51AFD7ED558CCD)
                h ^= h >> np.uint64(33)
        _, first, inverse = np.unique(h, return_index=True, return_inverse=True)
        table = np.ascontiguousarray(rows[first])
        if not np.array_equal(table[inverse], rows):
            raise RuntimeError("entry-row hash collision; use the inline form")
        inverse = inverse.astype(np.uint32)
        idx = inverse if was_inline else inverse[self.node_entry_row]
        out = dataclasses.replace(self, node_entry=None, node_entry_row=pin(np.ascontiguousarray(idx)),
                                  entry_rows=pin(np.ascontiguousarray(table[:, :es])))
        if fold_ph:
            out = dataclasses.replace(out, node_placeholders=None,
                                      row_placeholders=pin(np.ascontiguousarray(table[:, es:es + nph])))
        if fold_rnn:
            out = dataclasses.replace(out, node_rnn_id=None, row_rnn_id=pin(np.ascontiguousarray(table[:, -1])))
        return out

    def h2d_bytes(self) -> int:
        """Bytes jpp_b200_upload_batch copies to the device for this batch."""
        arrs = (self.sent_boundary_off, self.sent_codepoint_off, self.boundary_node_off, self.boundary_end_off,
                self.node_info, self.node_entry, self.node_placeholders, self.node_rnn_id, self.ends, self.codepoints,
                self.node_entry_row, self.entry_rows, self.row_placeholders, self.row_rnn_id, self.sent_extra_row_off)
        return int(sum(a.nbytes for a in arrs if a is not None))

    def resident_form(self, pin=None) -> tuple["StagedBatch", "EntryTable"]:
        """The same batch in the resident form of jpp_b200_batch (dictionary rows on the device, only the
        sentences' extra nodes bring rows) + the table of the dictionary entries it uses.  Built from an
        inline batch: a C++ stager makes the table from the dictionary once and never touches
        fillEntryBuffer for a dictionary node (binding/)."""
        assert self.node_entry is not None, "start from the inline form"
        pin = pin or (lambda a: a)
        ep = self.node_info["entry_ptr"]
        BOS, EOS = -(1 << 31), -(1 << 31) + 2
        dic = ep >= 0
        # dictionary table: one row per distinct entry_ptr
        uniq, first = np.unique(ep[dic], return_index=True)
        dic_idx = np.nonzero(dic)[0][first]
        table = EntryTable(np.ascontiguousarray(uniq.astype(np.int32)), np.ascontiguousarray(self.node_entry[dic_idx]),
                           None if self.node_rnn_id is None else np.ascontiguousarray(self.node_rnn_id[dic_idx]))
        # a function of the entry: check it
        inv = np.searchsorted(uniq, ep[dic])
        assert np.array_equal(table.rows[inv], self.node_entry[dic]), "entry buffer is not a function of entry_ptr"
        if self.node_rnn_id is not None:
            assert np.array_equal(table.rnn_id[inv], self.node_rnn_id[dic]), "rnn id is not a function of entry_ptr"
        if self.node_placeholders is not None:
            assert not self.node_placeholders[dic].any(), "dictionary nodes carry placeholders"
        # extra nodes: row index ~entry_ptr inside the sentence
        extra = (~dic) & (ep != BOS) & (ep != EOS)
        node_sent = np.searchsorted(self.sent_boundary_off, np.searchsorted(self.boundary_node_off, np.arange(len(ep)), side="right") - 1,
                                    side="right") - 1
        xi = (~ep[extra]).astype(np.int64)
        xs = node_sent[extra]
        cnt = np.zeros(self.num_sentences, np.int64)
        np.maximum.at(cnt, xs, xi + 1)
        off = np.zeros(self.num_sentences + 1, np.uint32)
        off[1:] = np.cumsum(cnt)
        nrows = int(off[-1])
        es = self.node_entry.shape[1]
        rows = np.zeros((max(1, nrows), es), np.int32)
        rows[off[xs] + xi] = self.node_entry[extra]
        ph = rid = None
        if self.node_placeholders is not None:
            ph = np.zeros((max(1, nrows), self.node_placeholders.shape[1]), np.int32)
            ph[off[xs] + xi] = self.node_placeholders[extra]
        if self.node_rnn_id is not None:
            rid = np.zeros(max(1, nrows), np.int32)
            rid[off[xs] + xi] = self.node_rnn_id[extra]
            # BOS / EOS nodes resolve to id 0 on the device
            assert not self.node_rnn_id[(ep == BOS) | (ep == EOS)].any()
        out = dataclasses.replace(self, node_entry=None, node_placeholders=None, node_rnn_id=None, node_entry_row=None,
                                  entry_rows=pin(rows[:max(1, nrows)]), row_placeholders=None if ph is None else pin(ph),
                                  row_rnn_id=None if rid is None else pin(rid), sent_extra_row_off=pin(off))
        return out, table


@dataclass
class EntryTable:
    """Pre-decoded dictionary entries for jpp_b200_add_entry_rows."""
    entry_ptr: np.ndarray        # [rows] i32, EntryPtr::rawValue() >= 0
    rows: np.ndarray             # [rows][entry_size] i32
    rnn_id: np.ndarray | None    # [rows] i32

    def as_struct(self) -> EntryTableStruct:
        return EntryTableStruct(len(self.entry_ptr), _ptr(self.entry_ptr), _ptr(self.rows), _ptr(self.rnn_id))


@dataclass
class ScoreResult:
    """Post-state of computeScores (jpp_b200_result)."""
    node_beam: np.ndarray          # [nodes][beam] BEAM_EL_DT
    node_beam_scores: np.ndarray   # [nodes][beam][S] f32
    node_patterns: np.ndarray      # [nodes][stored] u64
    node_t0_score: np.ndarray      # [nodes] f32
    best_path_len: np.ndarray | None = None     # [sentences] u32
    best_path_nodes: np.ndarray | None = None   # [boundaries] u32, sentence s at sent_boundary_off[s]
    sentence_status: np.ndarray | None = None   # [sentences] u8, 0 = scored (jpp_b200_sentence_status bits)

    @staticmethod
    def allocate(num_nodes: int, beam: int, num_scorers: int, stored: int, num_sentences: int | None = None,
                 num_boundaries: int = 0) -> "ScoreResult":
        r = ScoreResult(np.zeros((num_nodes, beam), BEAM_EL_DT),
                        np.zeros((num_nodes, beam, num_scorers), np.float32),
                        np.zeros((num_nodes, stored), np.uint64), np.zeros(num_nodes, np.float32))
        if num_sentences is not None:
            r.best_path_len = np.zeros(num_sentences, np.uint32)
            r.best_path_nodes = np.zeros(num_boundaries, np.uint32)
            r.sentence_status = np.zeros(num_sentences, np.uint8)
        return r

    def as_struct(self) -> Result:
        return Result(_ptr(self.node_beam), _ptr(self.node_beam_scores), _ptr(self.node_patterns),
                      _ptr(self.node_t0_score), _ptr(self.best_path_len), _ptr(self.best_path_nodes),
                      _ptr(self.sentence_status))

    def paths(self, batch: "StagedBatch") -> list[list[int]]:
        """1-best node indices per sentence from the device-resolved paths."""
        off = batch.sent_boundary_off
        return [self.best_path_nodes[int(off[s]):int(off[s]) + int(self.best_path_len[s])].tolist()
                for s in range(batch.num_sentences)]


def best_paths(batch: StagedBatch, res: ScoreResult) -> list[list[tuple[int, int]]]:
    """1-best path of every sentence as [(boundary, position), ...] from BOS to
    EOS: what AnalysisPath::fillIn walks (reference analysis_result.cc:25-76)."""
    out = []
    sbo, bno, beo = batch.sent_boundary_off, batch.boundary_node_off, batch.boundary_end_off
    for s in range(batch.num_sentences):
        b0 = int(sbo[s])
        nb = int(sbo[s + 1]) - b0
        path: list[tuple[int, int]] = []
        bnd, pos, slot = nb - 1, 0, 0
        while bnd >= 2:
            node = int(bno[b0 + bnd]) + pos
            el = res.node_beam[node, slot]
            if el["left"] == 0xFFFF:
                path = []
                break
            path.append((bnd, pos))
            ptr = batch.ends[int(beo[b0 + bnd]) + int(el["left"])]
            slot = int(el["prev_beam"])
            bnd, pos = int(ptr["boundary"]), int(ptr["position"])
        out.append(path[::-1])
    return out
