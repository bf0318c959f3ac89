"""Vocabulary facts the fused decode step needs, extracted from the reference `Tokenizer`.

The reference logits processors read these attributes of `osuT5.osuT5.tokenizer.Tokenizer`
(logit_processors.py:13-33,89-104,137-141; server.py:72-80): `event_start/event_end[EventType.*]`, `context_sos`,
`context_eos`, `sos_id/eos_id/pad_id`, `vocab_size_out`.  `TokenLayout.from_tokenizer` duck-types on those (enum keys
are matched by their `.value` string, so the reference's enums are never imported here); `from_json` loads the
committed v29 layout (`tests/golden/tokenizer_v29.json`) for boxes where the reference package is absent.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Dict, List, Optional

MILISECONDS_PER_STEP = 10  # server.py:22 (sic)

# osuT5/osuT5/dataset/data_utils.py:60-78 (TIMED_EVENTS), by EventType.value
TIMED_EVENT_VALUES = [
    "circle", "spinner", "spinner_end", "slider_head", "last_anchor", "slider_end", "beat", "measure",
    "timing_point", "kiai", "hold_note", "hold_note_end", "drumroll", "drumroll_end", "denden", "denden_end",
    "scroll_speed_change",
]


def _by_value(d) -> Dict[str, int]:
    return {(k.value if hasattr(k, "value") else str(k)): int(v) for k, v in d.items()}


@dataclasses.dataclass
class TokenLayout:
    vocab_size_out: int
    vocab_size_in: int
    pad_id: int
    sos_id: int
    eos_id: int
    event_start: Dict[str, int]
    event_end: Dict[str, int]
    context_sos: Dict[str, int]
    context_eos: Dict[str, int]

    # ---- construction -------------------------------------------------------------------------------------------
    @classmethod
    def from_tokenizer(cls, tok) -> "TokenLayout":
        if isinstance(tok, TokenLayout):
            return tok
        return cls(
            vocab_size_out=int(tok.vocab_size_out), vocab_size_in=int(tok.vocab_size_in),
            pad_id=int(tok.pad_id), sos_id=int(tok.sos_id), eos_id=int(tok.eos_id),
            event_start=_by_value(tok.event_start), event_end=_by_value(tok.event_end),
            context_sos=_by_value(getattr(tok, "context_sos", {})), context_eos=_by_value(getattr(tok, "context_eos", {})),
            time_shift_min_value=int({(k.value if hasattr(k, "value") else str(k)): v
                                      for k, v in tok.event_range.items()}["t"].min_value),
        )

    @classmethod
    def from_json(cls, path: str) -> "TokenLayout":
        with open(path) as f:
            return cls(**json.load(f))

    def to_json(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump(dataclasses.asdict(self), f, indent=1, sort_keys=True)

    # ---- derived sets (same derivations as the reference processors) ---------------------------------------------
    @property
    def time_shift_start(self) -> int:
        return self.event_start["t"]

    @property
    def time_shift_end(self) -> int:
        return self.event_end["t"]

    def sos_ids(self) -> List[int]:
        """logit_processors.py:141: [sos_id] + context_sos.values()"""
        return [self.sos_id] + list(self.context_sos.values())

    def lookback_eos_ids(self) -> List[int]:
        """logit_processors.py:97: [eos_id] + every context_eos"""
        return [self.eos_id] + list(self.context_eos.values())

    def timed_token_ids(self) -> List[int]:
        """logit_processors.py:100-104"""
        out: List[int] = []
        for v in TIMED_EVENT_VALUES:
            if v in self.event_start:
                out.extend(range(self.event_start[v], self.event_end[v]))
        return out

    def beat_type_tokens(self) -> List[int]:
        """logit_processors.py:13-20"""
        r = [self.event_start["beat"], self.event_start["measure"]]
        if "timing_point" in self.event_start:
            r.append(self.event_start["timing_point"])
        return r

    def mania_type_tokens(self) -> List[int]:
        """logit_processors.py:23-28"""
        if "hold_note_end" not in self.event_start:
            return []
        return [self.event_start["circle"], self.event_start["hold_note"], self.event_start["hold_note_end"]]

    def scroll_speed_tokens(self) -> List[int]:
        """logit_processors.py:31-33"""
        if "scroll_speed" not in self.event_start:
            return []
        return list(range(self.event_start["scroll_speed"], self.event_end["scroll_speed"]))

    def lookback_end(self, lookback_ms: float) -> int:
        """logit_processors.py:92: tokenizer.encode(Event(TIME_SHIFT, int(lookback/10))) = start + value - min_value.
        v29 has min_time_shift = 0 (tokenizer.py:86), so this is time_shift_start + int(lookback/10)."""
        return self.time_shift_start + int(lookback_ms / MILISECONDS_PER_STEP) - self.time_shift_min_value

    time_shift_min_value: int = 0

    def eos_token_ids(self, lookback_time: float = 0.0, lookahead_time: float = 0.0,
                      context_type: Optional[str] = None) -> List[int]:
        """server.py:72-80 `get_eos_token_id`."""
        ids = [self.eos_id]
        if context_type is not None:
            ct = context_type.value if hasattr(context_type, "value") else str(context_type)
            if ct in self.context_eos:
                ids.append(self.context_eos[ct])
        if lookback_time > 0:
            ids.extend(range(self.time_shift_start, self.time_shift_start + int(lookback_time / MILISECONDS_PER_STEP)))
        if lookahead_time > 0:
            ids.extend(range(self.time_shift_end - int(lookahead_time / MILISECONDS_PER_STEP), self.time_shift_end))
        return ids
