"""Token ids -> text.  Mirrors get_text's post-processing (marie/document/trocr_ocr_processor.py:155-178):
strip the fairseq specials, map ids through the dictionary, undo GPT-2 byte-level BPE
(GPT2BPEEnhancedSpace.decode, marie/models/unilm/trocr/bpe.py:59-67).

The real dictionary / BPE files (gpt2_with_mask.dict.txt, encoder.json, vocab.bpe) are downloaded assets that are not
part of the reference tree; point Gpt2Detokenizer at them when available.  SyntheticDetokenizer is the deterministic
stand-in used with random-init weights (same class on the device path and in the oracle, so text parity == id parity).
"""
import json
import os

SPECIALS = (0, 1, 2, 3)   # <s>, <pad>, </s>, <unk>


class _Pieces(dict):
    """token id -> text piece, filled on demand (decode is one dict lookup per token)"""

    def __missing__(self, t):
        k = int(t) - 4
        p = ((" " if k % 7 == 0 else "") + chr(97 + k % 26) + (chr(97 + (k // 26) % 26) if k % 3 else "")
             + (chr(97 + (k // 676) % 26) if k % 5 == 0 else ""))
        self[t] = p
        return p


class SyntheticDetokenizer:
    """id -> 1-3 letters; ids whose slot is 0 mod 7 start a new word (leading space), like GPT-2's 'Ġ' tokens."""

    def __init__(self):
        self._piece = _Pieces({t: "" for t in SPECIALS})

    def decode(self, ids):
        return "".join(map(self._piece.__getitem__, ids)).lstrip(" ")


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


class Gpt2Detokenizer:
    """get_text's tail (trocr_ocr_processor.py:162-176) without fairseq:
      fairseq Dictionary: ids 0-3 = <s> <pad> </s> <unk>, then one symbol per line of dict.txt ('<symbol> <count>';
        gpt2_with_mask.dict.txt lists GPT-2 ids as decimal strings and ends with '<mask>');
      Dictionary.string(hypo_tokens, extra_symbols_to_ignore={eos}): symbols joined by ' ', eos and bos dropped, <unk>
        and <pad> kept as their strings;
      GPT2BPEEnhancedSpace.decode (marie/models/unilm/trocr/bpe.py:59-67, INSERT_OR_REPLACE = 0): numeric symbols ->
        GPT-2 vocabulary strings (encoder.json inverted), '<unk>' / '<mask>' / '<s>' kept literally, byte-level
        decoding to UTF-8 (errors='replace'), then every '<s>' removed."""

    DICT_NAMES = ("gpt2_with_mask.dict.txt", "dict.txt")
    ENCODER_NAMES = ("encoder.json",)

    def __init__(self, dict_path, encoder_json_path):
        self.symbols = ["<s>", "<pad>", "</s>", "<unk>"]
        with open(dict_path, encoding="utf-8") as f:
            for line in f:
                if line.strip():
                    self.symbols.append(line.rstrip("\n").rsplit(" ", 1)[0])
        with open(encoder_json_path, encoding="utf-8") as f:
            enc = json.load(f)
        self.decoder = {v: k for k, v in enc.items()}
        self.byte_decoder = {v: k for k, v in _bytes_to_unicode().items()}

    @classmethod
    def locate(cls, *dirs):
        """first directory holding both files; FileNotFoundError naming what is missing otherwise"""
        for d in dirs:
            dp = next((os.path.join(d, n) for n in cls.DICT_NAMES if os.path.exists(os.path.join(d, n))), None)
            ep = next((os.path.join(d, n) for n in cls.ENCODER_NAMES if os.path.exists(os.path.join(d, n))), None)
            if dp and ep:
                return cls(dp, ep)
        raise FileNotFoundError(
            f"GPT-2 BPE files not found ({' or '.join(cls.DICT_NAMES)} and {cls.ENCODER_NAMES[0]}) in any of {list(dirs)}: pass "
            "detokenizer=Gpt2Detokenizer(dict_path, encoder_json_path); without them token ids cannot be turned into text")

    def decode(self, ids):
        syms = [self.symbols[t] if t < len(self.symbols) else "<unk>" for t in (int(t) for t in ids) if t not in (0, 2)]
        text = "".join(s if s in ("<unk>", "<mask>", "<s>") else self.decoder.get(int(s), s) if s.lstrip("-").isdigit() else s
                       for s in syms)
        return bytearray(self.byte_decoder[c] for c in text if c in self.byte_decoder).decode("utf-8", errors="replace").replace("<s>", "")
