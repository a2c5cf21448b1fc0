"""Constants of the reference that the hot path needs (values from videollama2/constants.py:4-32)."""
# placeholder ids the tokenizer step writes into input_ids; the splice replaces each by its visual tokens
MODAL_INDEX_MAP = {"<image>": -200, "<video>": -201, "<audio>": -202}
DEFAULT_IMAGE_TOKEN, DEFAULT_VIDEO_TOKEN, DEFAULT_AUDIO_TOKEN = tuple(MODAL_INDEX_MAP)
IMAGE_TOKEN_INDEX, VIDEO_TOKEN_INDEX, AUDIO_TOKEN_INDEX = tuple(MODAL_INDEX_MAP.values())

IGNORE_INDEX = -100              # label value of positions that carry no loss (visual tokens, padding)

NUM_FRAMES, MAX_FRAMES, NUM_FRAMES_PER_SECOND = 8, 32, 1   # default / cap of sampled frames; fps-mode sampling rate
