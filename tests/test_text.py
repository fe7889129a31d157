"""Token inventory / tokenizer against the reference's own test vector (tests/test_tokenizer.py:8-13)."""
from forwardtacotron_b200.utils.text import NUM_PHONEMES, PAD_ID, Tokenizer, phonemes


def test_reference_tokenizer_vector():
    tok = Tokenizer()
    assert tok('_ abc{') == [0, 10, 36, 52, 57]          # '{' is not in the inventory and is dropped
    assert tok.decode([0, 10, 36, 52, 57]) == '_ abc'


def test_inventory_is_the_embedding_contract():
    assert NUM_PHONEMES == 135 and phonemes[PAD_ID] == '_' and len(set(phonemes)) == 135
