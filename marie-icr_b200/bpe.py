"""Token ids -> text.  Mirrors get_text's post-processing (marie/document/trocr_ocr_processor.py:155-178):
strip the fairseq specials, map ids through the dictionary, undo GPT-2 byte-level BPE
(GPT2BPEEnhancedSpace.decode, marie/models/unilm/trocr/bpe.py:59-67).

The real dictionary / BPE files (gpt2_with_mask.dict.txt, encoder.json, vocab.bpe) are downloaded assets that are not
part of the reference tree; point Gpt2Detokenizer at them when available.  SyntheticDetokenizer is the deterministic
stand-in used with random-init weights (same class on the device path and in the oracle, so text parity == id parity).
"""
import json

SPECIALS = (0, 1, 2, 3)   # <s>, <pad>, </s>, <unk>


class SyntheticDetokenizer:
    """id -> 1-3 letters; ids whose slot is 0 mod 7 start a new word (leading space), like GPT-2's 'Ġ' tokens."""

    def decode(self, ids):
        out = []
        for t in ids:
            t = int(t)
            if t in SPECIALS:
                continue
            k = t - 4
            s = chr(97 + k % 26) + (chr(97 + (k // 26) % 26) if k % 3 else "") + (chr(97 + (k // 676) % 26) if k % 5 == 0 else "")
            out.append((" " if k % 7 == 0 and out else "") + s)
        return "".join(out)


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
    """fairseq dictionary (dict.txt: '<gpt2 id> <count>' per line, ids offset by the 4 specials) + GPT-2 encoder.json."""

    def __init__(self, dict_path, encoder_json_path):
        self.symbols = ["<s>", "<pad>", "</s>", "<unk>"]
        with open(dict_path, encoding="utf-8") as f:
            for line in f:
                if line.strip():
                    self.symbols.append(line.rsplit(" ", 1)[0])
        with open(encoder_json_path, encoding="utf-8") as f:
            enc = json.load(f)
        self.decoder = {v: k for k, v in enc.items()}
        self.byte_decoder = {v: k for k, v in _bytes_to_unicode().items()}

    def decode(self, ids):
        toks = [self.symbols[int(t)] for t in ids if int(t) not in SPECIALS and int(t) < len(self.symbols)]
        text = "".join(self.decoder.get(int(t), t) if t.lstrip("-").isdigit() else t for t in toks if t != "<mask>")
        return bytearray(self.byte_decoder[c] for c in text if c in self.byte_decoder).decode("utf-8", errors="replace")
