"""Token inventory and tokenizer of the reference (utils/text/symbols.py:8-23, utils/text/tokenizer.py:6-16).

The inventory is interface data, not code: id = position in this list, pad '_' = id 0, 135 entries = the embedding
rows of every checkpoint.  The cleaners / espeak phonemizer (utils/text/cleaners.py) are a CPU text front-end outside
this package: callers pass phonemised text (what ``Cleaner`` produces upstream)."""
from __future__ import annotations

from typing import List

_pad = '_'
_punctuation = '!\'(),.:;? '
_special = '-'
_vowels = 'iyɨʉɯuɪʏʊeøɘəɵɤoɛœɜɞʌɔæɐaɶɑɒᵻ'
_non_pulmonic_consonants = 'ʘɓǀɗǃʄǂɠǁʛ'
_pulmonic_consonants = 'pbtdʈɖcɟkɡqɢʔɴŋɲɳnɱmʙrʀⱱɾɽɸβfvθðszʃʒʂʐçʝxɣχʁħʕhɦɬɮʋɹɻjɰlɭʎʟ'
_suprasegmentals = 'ˈˌːˑ'
_other_symbols = 'ʍwɥʜʢʡɕʑɺɧ'
_diacritics = 'ɚ˞ɫ'
_extra_phons = ['g', 'ɝ', '̃', '̍', '̥', '̩', '̯', '͡']

phonemes: List[str] = list(_pad + _punctuation + _special + _vowels + _non_pulmonic_consonants + _pulmonic_consonants
                           + _suprasegmentals + _other_symbols + _diacritics) + _extra_phons
NUM_PHONEMES = len(phonemes)
PAD_ID = 0
assert NUM_PHONEMES == 135 and phonemes[PAD_ID] == '_'


class Tokenizer:
    """text -> ids; symbols outside the inventory are dropped (utils/text/tokenizer.py:13)."""

    def __init__(self) -> None:
        self.symbol_to_id = {s: i for i, s in enumerate(phonemes)}
        self.id_to_symbol = dict(enumerate(phonemes))

    def __call__(self, text: str) -> List[int]:
        return [self.symbol_to_id[t] for t in text if t in self.symbol_to_id]

    def decode(self, sequence: List[int]) -> str:
        return ''.join(self.id_to_symbol[s] for s in sequence if s in self.id_to_symbol)
