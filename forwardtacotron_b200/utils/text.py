"""Size of the reference's phoneme inventory (utils/text/symbols.py:21-23): pad '_' (id 0) +
punctuation + IPA symbols = 135 entries.  Only the count reaches the hot path (embedding rows);
the text front-end itself (cleaners, phonemizer) is outside the scope of this package."""
NUM_PHONEMES = 135
PAD_ID = 0
