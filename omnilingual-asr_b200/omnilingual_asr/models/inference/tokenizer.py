"""id -> text for the CTC head (a17).

omniASR uses a 9812-entry SentencePiece character vocabulary (`omniASR_tokenizer`); the file is not
available offline, so the table is injected: a list of pieces, a SentencePiece model path (loaded with the
`sentencepiece` package), or the deterministic synthetic table used by tests and benchmarks.
Ids 0..3 are specials (blank/pad, bos, eos, unk) and are dropped, as upstream's
`skip_special_tokens=True` does.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import numpy as np

WORD_BOUNDARY = "▁"  # SentencePiece whitespace marker
N_SPECIALS = 4


class CtcVocabulary:
    def __init__(self, pieces: Sequence[str], n_specials: int = N_SPECIALS):
        self.pieces = list(pieces)
        self.n_specials = n_specials
        # specials map to the empty string in the lookup table: decode is one fancy index + one join (a 9.5 h
        # recording is 1.2 M tokens; a Python loop over them cost 0.2 s, a third of the 8-GPU wall time)
        self._table = np.array([("" if i < n_specials else p) for i, p in enumerate(self.pieces)], dtype=object)

    def __len__(self) -> int:
        return len(self.pieces)

    @classmethod
    def synthetic(cls, vocab_size: int) -> "CtcVocabulary":
        """Deterministic stand-in: specials, the word boundary, a-z, then unique code points."""
        pieces = ["<blank>", "<s>", "</s>", "<unk>", WORD_BOUNDARY]
        pieces += [chr(ord("a") + i) for i in range(26)]
        cp = 0x4E00  # CJK block: printable, one code point per id
        while len(pieces) < vocab_size:
            pieces.append(chr(cp))
            cp += 1
        return cls(pieces[:vocab_size])

    @classmethod
    def from_sentencepiece(cls, model_path: str) -> "CtcVocabulary":
        import sentencepiece as spm
        sp = spm.SentencePieceProcessor(model_file=model_path)
        return cls([sp.id_to_piece(i) for i in range(sp.get_piece_size())])

    def decode(self, ids: Iterable[int]) -> str:
        a = np.asarray(ids if isinstance(ids, np.ndarray) else list(ids), dtype=np.int64)
        if a.size == 0:
            return ""
        a = a[(a >= self.n_specials) & (a < len(self.pieces))]     # specials and out-of-table ids are dropped
        return "".join(self._table[a].tolist()).replace(WORD_BOUNDARY, " ").strip()

    def words_with_frames(self, ids: Sequence[int], frames: Sequence[int]) -> List[Tuple[str, int, int]]:
        """Group tokens into words at the boundary marker -> (word, first_frame, last_frame)."""
        words: List[Tuple[str, int, int]] = []
        cur, f0, f1 = "", -1, -1
        for i, f in zip(ids, frames):
            i = int(i)
            if i < self.n_specials or i >= len(self.pieces):
                continue
            piece = self.pieces[i]
            if piece.startswith(WORD_BOUNDARY):
                if cur:
                    words.append((cur, f0, f1))
                cur, f0, f1 = "", -1, -1
                piece = piece[len(WORD_BOUNDARY):]
                if not piece:
                    continue
            if not cur:
                f0 = int(f)
            cur += piece
            f1 = int(f)
        if cur:
            words.append((cur, f0, f1))
        return words
