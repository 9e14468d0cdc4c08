"""Rule-based Russian grapheme->phoneme fallback for out-of-dictionary words.

Behavioural mirror of `vosk_tts.g2p.convert` (/root/reference/vosk_tts/g2p.py:84-107): a stress mark `+` precedes the
stressed vowel; consonants that have a soft/hard pair are palatalised (suffix `j`) before a soft letter; the iotated
vowels я ю е ё emit a leading `j` at the start of a syllable; vowels carry the stress digit (0/1).  Host-side string
processing (microseconds per word) -- it feeds the phoneme ids that the CUDA engine consumes.
"""

_PAIRED = {"б": "b", "в": "v", "г": "g", "Г": "g", "д": "d", "з": "z", "к": "k", "л": "l", "м": "m", "н": "n", "п": "p",
           "р": "r", "с": "s", "т": "t", "ф": "f", "х": "h"}
_UNPAIRED = {"ж": "zh", "ц": "c", "ч": "ch", "ш": "sh", "щ": "sch", "й": "j"}
_VOWEL = {"а": "a", "я": "a", "у": "u", "ю": "u", "о": "o", "ё": "o", "э": "e", "е": "e", "и": "i", "ы": "y"}
_SOFTENERS = frozenset("яёюиье")
_SYLLABLE_START = frozenset("#ъьаяоёуюэеиы-")
_IOTATED = frozenset("яюеё")
_SILENT = frozenset(["#", "+", "-", "ь", "ъ"])


def convert(stressword):
    # 1. attach the stress flag carried by '+' to the letter that follows it
    letters = []
    stressed = 0
    for ch in "#" + stressword + "#":
        if ch == "+":
            stressed = 1
            continue
        letters.append([ch, stressed])
        stressed = 0
    # 2. consonants (the closing '#' is never rewritten); a rewritten symbol is no longer a letter for step 3
    symbols = [l[0] for l in letters]
    for i in range(len(letters) - 1):
        ch = letters[i][0]
        if ch in _PAIRED:
            symbols[i] = _PAIRED[ch] + ("j" if letters[i + 1][0] in _SOFTENERS else "")
            letters[i][1] = 0
        if ch in _UNPAIRED:
            symbols[i] = _UNPAIRED[ch]
            letters[i][1] = 0
    # 3. vowels, iotation after a syllable start (decided on the already rewritten previous symbol, as the reference does)
    out = []
    prev = ""
    for sym, (ch, st) in zip(symbols, letters):
        if prev in _SYLLABLE_START and sym in _IOTATED:
            out.append("j")
        if sym in _VOWEL:
            out.append(_VOWEL[sym] + str(st))
        else:
            out.append(sym)
        prev = sym
    return " ".join(p for p in out if p not in _SILENT)
