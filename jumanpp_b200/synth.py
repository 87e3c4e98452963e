"""Seeded synthetic workload for the Juman++ lattice-scoring path.

No Jumandic dictionary, trained model or RNN model ships with the reference
(reference README.md:40-41, test/CMakeLists.txt:1-6), so every artefact the
hot path needs is manufactured here from a seed:

* ``write_dictionary``  - a Jumandic-format CSV lexicon (the column layout the
  reference's spec imports: src/jumandic/shared/jumandic_spec.cc:31-45;
  8 leading UNK template rows are required by the unk makers, :68-103).
* ``write_corpus``      - UTF-8 sentences, one per line: mostly concatenations
  of dictionary surfaces (Zipf), some random kanji/kana runs, a few
  ASCII/digit runs so the unknown-word makers fire.
* ``make_weights``      - the hashed perceptron table ``float[2**k]``
  (reference src/core/analysis/perceptron.cc:81-124 loads exactly this).
  Generated from a counter-based integer hash so that the CPU oracle, the
  reference harness and the CUDA side can all materialise identical tables
  without shipping 256 MiB files (same formula in oracle/jpp_oracle.c,
  oracle/ref_harness.cc and jumanpp_b200/csrc/jpp_b200_kernels.cu).
* ``write_rnn_model``   - a faster-rnnlm ``.nnet`` v6 file + vocabulary in the
  layout the reference parses (src/rnn/mikolov_rnn.cc:16-76,163-215).

Everything is deterministic in (seed, sizes); numpy's legacy ``RandomState``
is used because its streams are stable across numpy versions.
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

# --------------------------------------------------------------------------
# character pools
# --------------------------------------------------------------------------
HIRAGANA = [chr(c) for c in range(0x3042, 0x3094) if chr(c) not in "ゐゑ"]
KATAKANA = [chr(c) for c in range(0x30A2, 0x30F4) if chr(c) not in "ヰヱ"]
# 2000 kanji from the CJK unified block, fixed stride (not frequency ordered;
# the Zipf ranks below give them a frequency order)
KANJI = [chr(0x4E00 + i * 7) for i in range(2000)]
ASCII_LETTERS = list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
DIGITS = list("0123456789０１２３４５６７８９")
SYMBOLS = list("、。・「」！？…―（）")

UNK_ROWS = [
    # surface, pos, subpos, conjform, conjtype, baseform, reading, canonic, features
    ("UNK_SYM", "未定義語", "その他", "*", "*", "UNK", "UNK", "*", "品詞推定:特殊 未知語:その他"),
    ("UNK_KATA", "未定義語", "カタカナ", "*", "*", "UNK", "UNK", "*", "品詞推定:名詞 未知語:カタカナ"),
    ("UNK_KANJI", "未定義語", "その他", "*", "*", "UNK", "UNK", "*", "品詞推定:名詞 未知語:漢字"),
    ("UNK_HIRA", "未定義語", "その他", "*", "*", "UNK", "UNK", "*", "品詞推定:名詞 未知語:ひらがな"),
    ("UNK_ALPH", "未定義語", "アルファベット", "*", "*", "UNK", "UNK", "*", "品詞推定:名詞 未知語:ロマ字"),
    ("UNK_DIGIT", "名詞", "数詞", "*", "*", "UNK", "UNK", "*", "カテゴリ:数量 未知語:数字"),
    ("UNK_ONOMATOPEA", "副詞", "*", "*", "*", "UNK", "UNK", "*", "自動認識 未知語:擬音語擬態語"),
    ("UNK_UNK_CHAR_CAT", "未定義語", "その他", "*", "*", "UNK", "UNK", "*", "品詞推定:特殊 未知語:未対応文字種"),
]

# closed-class words: exercise the auxWord / lexicalized feature branches
# (jumandic_spec.cc:47-63 matches pos in {助詞, 助動詞, 判定詞} and the
# lexicalized (baseform,pos,subpos,conjtype) list)
FUNCTION_WORDS = [
    ("は", "助詞", "副助詞", "*", "*"), ("が", "助詞", "格助詞", "*", "*"),
    ("が", "助詞", "接続助詞", "*", "*"), ("を", "助詞", "格助詞", "*", "*"),
    ("に", "助詞", "格助詞", "*", "*"), ("で", "助詞", "格助詞", "*", "*"),
    ("と", "助詞", "格助詞", "*", "*"), ("も", "助詞", "副助詞", "*", "*"),
    ("の", "助詞", "接続助詞", "*", "*"), ("の", "助詞", "格助詞", "*", "*"),
    ("へ", "助詞", "格助詞", "*", "*"), ("から", "助詞", "格助詞", "*", "*"),
    ("まで", "助詞", "格助詞", "*", "*"), ("より", "助詞", "格助詞", "*", "*"),
    ("や", "助詞", "接続助詞", "*", "*"), ("など", "助詞", "副助詞", "*", "*"),
    ("か", "助詞", "終助詞", "*", "*"), ("ね", "助詞", "終助詞", "*", "*"),
    ("よ", "助詞", "終助詞", "*", "*"), ("ので", "助詞", "接続助詞", "*", "*"),
    ("だ", "判定詞", "*", "基本形", "判定詞"), ("です", "判定詞", "*", "デス列基本形", "判定詞"),
    ("で", "判定詞", "*", "タ系連用テ形", "判定詞"), ("だった", "判定詞", "*", "ダ列タ形", "判定詞"),
    ("らしい", "助動詞", "*", "基本形", "イ形容詞イ段"), ("のだ", "助動詞", "*", "基本形", "ナ形容詞"),
    ("だろう", "助動詞", "*", "基本形", "助動詞だろう型"), ("そうだ", "助動詞", "*", "基本形", "ナ形容詞"),
    ("、", "特殊", "読点", "*", "*"), ("。", "特殊", "句点", "*", "*"),
    ("・", "特殊", "記号", "*", "*"), ("「", "特殊", "括弧始", "*", "*"),
    ("」", "特殊", "括弧終", "*", "*"), ("！", "特殊", "記号", "*", "*"),
    ("こと", "名詞", "形式名詞", "*", "*"), ("の", "名詞", "形式名詞", "*", "*"),
    ("日本", "名詞", "地名", "*", "*"), ("者", "接尾辞", "名詞性名詞接尾辞", "*", "*"),
    ("日", "接尾辞", "名詞性名詞助数辞", "*", "*"), ("年", "接尾辞", "名詞性名詞助数辞", "*", "*"),
    ("する", "動詞", "*", "基本形", "サ変動詞"), ("し", "動詞", "*", "基本連用形", "サ変動詞"),
    ("いる", "接尾辞", "動詞性接尾辞", "基本形", "母音動詞"), ("い", "接尾辞", "動詞性接尾辞", "基本連用形", "母音動詞"),
    ("ない", "接尾辞", "形容詞性述語接尾辞", "基本形", "イ形容詞アウオ段"),
    ("れる", "接尾辞", "動詞性接尾辞", "基本形", "母音動詞"),
    ("なる", "動詞", "*", "基本形", "子音動詞ラ行"), ("ある", "動詞", "*", "基本形", "子音動詞ラ行"),
    ("いう", "動詞", "*", "基本形", "子音動詞ワ行"), ("一", "名詞", "数詞", "*", "*"),
]

# open-class categories: (pos, subpos, weight, [(conjform, ending)...] or None, conjtype)
OPEN_CLASSES = [
    ("名詞", "普通名詞", 38.0, None, "*"),
    ("名詞", "サ変名詞", 18.0, None, "*"),
    ("名詞", "地名", 4.0, None, "*"),
    ("名詞", "人名", 3.0, None, "*"),
    ("名詞", "組織名", 1.5, None, "*"),
    ("名詞", "時相名詞", 2.5, None, "*"),
    ("名詞", "副詞的名詞", 1.0, None, "*"),
    ("副詞", "*", 3.0, None, "*"),
    ("連体詞", "*", 0.5, None, "*"),
    ("接続詞", "*", 0.6, None, "*"),
    ("感動詞", "*", 0.4, None, "*"),
    ("指示詞", "名詞形態指示詞", 0.5, None, "*"),
    ("接頭辞", "名詞接頭辞", 1.2, None, "*"),
    ("接尾辞", "名詞性名詞接尾辞", 2.5, None, "*"),
    ("接尾辞", "名詞性名詞助数辞", 1.5, None, "*"),
    ("動詞", "*", 9.0, [("基本形", "る"), ("未然形", ""), ("基本連用形", ""), ("タ形", "た"), ("タ系連用テ形", "て"), ("命令形", "ろ")], "母音動詞"),
    ("動詞", "*", 5.0, [("基本形", "る"), ("未然形", "ら"), ("基本連用形", "り"), ("タ形", "った"), ("タ系連用テ形", "って"), ("意志形", "ろう")], "子音動詞ラ行"),
    ("動詞", "*", 2.5, [("基本形", "く"), ("未然形", "か"), ("基本連用形", "き"), ("タ形", "いた"), ("タ系連用テ形", "いて")], "子音動詞カ行"),
    ("動詞", "*", 1.5, [("基本形", "う"), ("未然形", "わ"), ("基本連用形", "い"), ("タ形", "った"), ("タ系連用テ形", "って")], "子音動詞ワ行"),
    ("形容詞", "*", 3.0, [("基本形", "い"), ("語幹", ""), ("基本連用形", "く"), ("タ形", "かった"), ("文語基本形", "し")], "イ形容詞アウオ段"),
    ("形容詞", "*", 2.0, [("基本形", "だ"), ("語幹", ""), ("ダ列基本連体形", "な"), ("ダ列タ系連用テ形", "で"), ("ダ列基本連用形", "に")], "ナ形容詞"),
]

FEATURE_POOL = ["NIL", "NIL", "NIL", "NIL", "内容語", "準内容語", "カテゴリ:抽象物",
                "カテゴリ:人", "カテゴリ:場所-施設", "ドメイン:政治", "ドメイン:ビジネス",
                "自他動詞:他:x/y", "濁音化D", "連用形名詞化", "連用形名詞化 内容語",
                "付属動詞候補（基本）", "漢字読み:音", "漢字読み:訓"]


def _zipf_probs(n: int, s: float) -> np.ndarray:
    r = np.arange(1, n + 1, dtype=np.float64)
    p = r ** (-s)
    return p / p.sum()


class CharSampler:
    """Zipf(s=1.1) over a kanji / hiragana / katakana mix."""

    def __init__(self, rng: np.random.RandomState):
        self.rng = rng
        self.pools = [KANJI, HIRAGANA, KATAKANA]
        self.pool_p = np.array([0.55, 0.33, 0.12])
        self.probs = [_zipf_probs(len(p), 1.1) for p in self.pools]

    def word(self, length: int, script: int | None = None) -> str:
        if script is None:
            script = int(self.rng.choice(3, p=self.pool_p))
        pool = self.pools[script]
        idx = self.rng.choice(len(pool), size=length, p=self.probs[script])
        return "".join(pool[i] for i in idx)


def _reading(rng: np.random.RandomState, n: int) -> str:
    idx = rng.randint(0, len(HIRAGANA), size=max(1, n))
    return "".join(HIRAGANA[i] for i in idx)


def generate_dictionary_rows(n_rows: int, seed: int) -> list[tuple[str, ...]]:
    """Rows are (surface,pos,subpos,conjform,conjtype,baseform,reading,canonic,features)."""
    rng = np.random.RandomState(seed)
    cs = CharSampler(rng)
    rows: list[tuple[str, ...]] = list(UNK_ROWS)
    seen: set[tuple[str, ...]] = set()

    per_surface: dict[str, int] = {}

    def add(row: tuple[str, ...], cap: int = 4) -> None:
        # at most `cap` homographs per surface (a real lexicon has a handful;
        # without the cap the Zipf head characters collect hundreds of 1-char
        # entries and every lattice boundary explodes)
        key = row[:7]
        if key in seen or per_surface.get(row[0], 0) >= cap:
            return
        seen.add(key)
        per_surface[row[0]] = per_surface.get(row[0], 0) + 1
        rows.append(row)

    for surf, pos, sub, cform, ctype in FUNCTION_WORDS:
        base = surf
        if (pos, ctype) == ("判定詞", "判定詞"):
            base = "だ"
        elif ctype == "サ変動詞":
            base = "する"
        elif surf == "い" and ctype == "母音動詞":
            base = "いる"
        add((surf, pos, sub, cform, ctype, base, surf, "*", "NIL"), cap=8)

    cls_w = np.array([c[2] for c in OPEN_CLASSES])
    cls_p = cls_w / cls_w.sum()
    len_p = np.array([0.10, 0.42, 0.24, 0.14, 0.07, 0.03])  # 1..6 chars
    while len(rows) < n_rows:
        ci = int(rng.choice(len(OPEN_CLASSES), p=cls_p))
        pos, sub, _, conj, ctype = OPEN_CLASSES[ci]
        length = int(rng.choice(6, p=len_p)) + 1
        homographs = 1 + int(rng.rand() < 0.18) + int(rng.rand() < 0.05)
        feats = FEATURE_POOL[int(rng.randint(0, len(FEATURE_POOL)))]
        if conj is None:
            surf = cs.word(length)
            for h in range(homographs):
                if h == 0:
                    p2, s2 = pos, sub
                else:
                    cj = int(rng.choice(15, p=cls_p[:15] / cls_p[:15].sum()))
                    p2, s2 = OPEN_CLASSES[cj][0], OPEN_CLASSES[cj][1]
                rd = _reading(rng, length + 1)
                add((surf, p2, s2, "*", "*", surf, rd, f"{surf}/{rd}", feats))
        else:
            stem = cs.word(max(1, length - 1), script=0 if rng.rand() < 0.8 else 1)
            base = stem + conj[0][1]
            rd_stem = _reading(rng, length)
            for cform, ending in conj:
                surf = stem + ending
                if not surf:
                    continue
                add((surf, pos, sub, cform, ctype, base, rd_stem + ending,
                     f"{base}/{rd_stem}{conj[0][1]}", feats))
    return rows[:n_rows]


def write_dictionary(path: str | Path, n_rows: int = 200_000, seed: int = 20260923) -> list[str]:
    """Write the CSV and return the list of (non-UNK) surfaces in row order."""
    rows = generate_dictionary_rows(n_rows, seed)
    with open(path, "w", encoding="utf-8") as f:
        for r in rows:
            surf, pos, sub, cform, ctype, base, rd, canon, feats = r
            f.write(f"{surf},0,0,0,{pos},{sub},{cform},{ctype},{base},{rd},{canon},{feats}\n")
    return [r[0] for r in rows[len(UNK_ROWS):]]


def rnn_vocabulary(path_csv: str | Path, size: int) -> list[str]:
    """``</s>``, ``<unk>`` then ``baseform_pos`` keys in dictionary order
    (rnnFields = [baseform, pos], separator ``_``: reference
    src/core/analysis/rnn_scorer.h:19-28)."""
    words = ["</s>", "<unk>"]
    seen = set(words)
    with open(path_csv, encoding="utf-8") as f:
        for i, line in enumerate(f):
            if i < len(UNK_ROWS):
                continue
            c = line.rstrip("\n").split(",")
            key = f"{c[8]}_{c[4]}"
            if key not in seen:
                seen.add(key)
                words.append(key)
                if len(words) >= size:
                    break
    return words


_PARTICLES = [w[0] for w in FUNCTION_WORDS if w[1] in ("助詞", "判定詞", "助動詞")]


def generate_corpus(n_sent: int, min_len: int, max_len: int, seed: int,
                    surfaces: list[str]) -> list[str]:
    rng = np.random.RandomState(seed)
    cs = CharSampler(rng)
    n_surf = len(surfaces)
    # Zipf over dictionary rank, flattened so the long tail is visited
    surf_p = _zipf_probs(n_surf, 0.9)
    surf_cdf = np.cumsum(surf_p)
    out: list[str] = []
    for _ in range(n_sent):
        target = int(rng.randint(min_len, max_len + 1))
        parts: list[str] = []
        total = 0
        while total < target:
            u = rng.rand()
            if u < 0.70:
                w = surfaces[int(np.searchsorted(surf_cdf, rng.rand()))]
            elif u < 0.95:
                w = cs.word(int(rng.randint(1, 5)))
            elif u < 0.975:
                k = int(rng.randint(1, 6))
                w = "".join(ASCII_LETTERS[i] for i in rng.randint(0, len(ASCII_LETTERS), size=k))
            elif u < 0.99:
                k = int(rng.randint(1, 5))
                w = "".join(DIGITS[i] for i in rng.randint(0, len(DIGITS), size=k))
            else:
                w = SYMBOLS[int(rng.randint(0, len(SYMBOLS)))]
            parts.append(w)
            total += len(w)
            # Japanese alternates scripts: content words are mostly followed by a
            # hiragana particle / auxiliary, which keeps same-script runs short
            # (the unk makers enumerate every sub-span of a same-script run:
            # reference src/core/analysis/unk_nodes_creator.cc:19-65)
            if rng.rand() < 0.65:
                fw = _PARTICLES[int(rng.randint(0, len(_PARTICLES)))]
                parts.append(fw)
                total += len(fw)
        s = "".join(parts)[:target]
        out.append(s)
    return out


def write_corpus(path: str | Path, n_sent: int, min_len: int, max_len: int, seed: int,
                 surfaces: list[str]) -> None:
    sents = generate_corpus(n_sent, min_len, max_len, seed, surfaces)
    with open(path, "w", encoding="utf-8") as f:
        for s in sents:
            f.write(s + "\n")


# --------------------------------------------------------------------------
# counter-based weights: w[i] = scale * (irwin_hall4(mix(seed, i)) - 2)
# --------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def make_weights(log2_size: int, seed: int, scale: float = 0.5, offset: int = 0,
                 count: int | None = None) -> np.ndarray:
    """float32 table; element i depends only on (seed, i).

    value = scale * ((a + b + c + d) / 65536 - 2) * sqrt(3), a..d the four 16-bit
    lanes of splitmix64(seed * 2^40 + i): Irwin-Hall(4), unit variance after the
    sqrt(3) factor; computed in float32 exactly as the C / CUDA twins do.
    """
    n = (1 << log2_size) if count is None else count
    with np.errstate(over="ignore"):
        idx = np.arange(offset, offset + n, dtype=np.uint64)
        h = _splitmix64(idx + (np.uint64(seed) << np.uint64(40)))
    a = (h & np.uint64(0xFFFF)).astype(np.uint32)
    b = ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.uint32)
    c = ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.uint32)
    d = ((h >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.uint32)
    s = (a + b + c + d).astype(np.float32)  # exact: < 2^18
    v = s * np.float32(1.0 / 65536.0) - np.float32(2.0)
    factor = np.float32(float(np.float32(scale)) * 1.7320508075688772)
    return (v * factor).astype(np.float32)


# --------------------------------------------------------------------------
# RNN model files
# --------------------------------------------------------------------------
def write_rnn_model(prefix: str | Path, words: list[str], layer_size: int = 128,
                    maxent_order: int = 3, maxent_size: int = 1 << 22,
                    nce_lnz: float = 9.0, seed: int = 2) -> None:
    """``<prefix>`` = vocabulary (one word + count per line, space separated,
    first field is the word: mikolov_rnn.cc:171-178), ``<prefix>.nnet`` = header +
    embeddings[V][H] + nce embeddings[V][H] + matrix[H][H] + maxent[maxent_size]
    (mikolov_rnn.cc:29-76 header fields in file order, :196-209 blocks)."""
    prefix = Path(prefix)
    with open(prefix, "w", encoding="utf-8") as f:
        for i, w in enumerate(words):
            f.write(f"{w} {len(words) - i}\n")
    v = len(words)
    h = layer_size
    with open(str(prefix) + ".nnet", "wb") as f:
        f.write(struct.pack("<Q", 6 * 10000 + h))       # sizeVersion
        f.write(struct.pack("<Q", maxent_size))          # maxEntTableSize
        f.write(struct.pack("<I", maxent_order))
        f.write(struct.pack("<?", True))                 # useNce
        f.write(struct.pack("<f", nce_lnz))
        f.write(struct.pack("<?", False))                # reversedSentence
        f.write(b"sigmoid".ljust(64, b"\0"))             # layerType[64]
        f.write(struct.pack("<I", 1))                    # layerCount
        f.write(struct.pack("<I", 0))                    # hsArity
        make_weights(0, seed * 16 + 1, 0.1, count=v * h).tofile(f)       # embeddings
        make_weights(0, seed * 16 + 2, 0.1, count=v * h).tofile(f)       # nce embeddings
        make_weights(0, seed * 16 + 3, 0.1, count=h * h).tofile(f)       # recurrent matrix
        make_weights(0, seed * 16 + 4, 0.05, count=maxent_size).tofile(f)  # maxent
