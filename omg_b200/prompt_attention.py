"""Prompt-to-prompt controller for the fused attention path.

Keeps the reference's controller surface (src/prompt_attention/p2p_attention.py:140-147 `AttentionReplace(prompts,
num_steps, cross_replace_steps, self_replace_steps, width, height, local_blend=None, tokenizer=None, device=None,
dtype=None)`, attributes `cross_replace_alpha`, `mapper`, `num_self_replace`, `batch_size`, `num_att_layers`,
`cur_step`, `cur_att_layer`, `reset()`), but instead of editing a materialised probability tensor it hands the fused
kernels an *edit specification*:

  self-attention   base probabilities replace the edited image's  <=>  out_1 = softmax(Q_0 K_0^T) V_1
                   (p2p_attention.py:114-118,126,136) -> batch-row remap (q,k from image 0, v from image 1);
  cross-attention  P_1 <- (P_0 M) * alpha + (1 - alpha) * P_1   (p2p_attention.py:131-134,146-147)
                   <=>  out_1 = P_0 (M diag(alpha) V_1) + P_1 (diag(1 - alpha) V_1): two plain attention terms over
                   contexts mixed once per step by omg_ctx_mix.

The layer/step counters advance exactly like AttentionControl.__call__ (p2p_attention.py:28-40): one step per
`num_att_layers` attention calls, so `reset()` / `cur_step` behave as in the reference.
"""
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

MAX_WORDS = 77


def _word_token_spans(text: str, tokenizer) -> List[List[int]]:
    """Token positions (1-based: BOS is 0) of every whitespace-separated word, the way seq_aligner.get_word_inds
    (seq_aligner.py:5-23) assigns sub-word tokens to words: tokens are consumed until their stripped character count
    reaches the word's length."""
    words = text.split(" ")
    pieces = [tokenizer.decode([tid]).strip("#") for tid in tokenizer.encode(text)][1:-1]
    spans: List[List[int]] = [[] for _ in words]
    w, used = 0, 0
    for pos, piece in enumerate(pieces):
        if w >= len(words):
            break
        spans[w].append(pos + 1)
        used += len(piece)
        if used >= len(words[w]):
            w, used = w + 1, 0
    return spans


def replacement_mapper(prompts: Sequence[str], tokenizer, max_len: int = MAX_WORDS) -> torch.Tensor:
    """(len(prompts)-1, 77, 77) token-alignment matrices between prompt 0 and every edited prompt
    (seq_aligner.py:25-66).  Raises ValueError when word counts differ, like the reference (seq_aligner.py:28-30)."""
    base = prompts[0]
    out = []
    for other in prompts[1:]:
        wa, wb = base.split(" "), other.split(" ")
        if len(wa) != len(wb):
            raise ValueError("attention replacement edit can only be applied on prompts with the same length"
                             f" but prompt A has {len(wa)} words and prompt B has {len(wb)} words.")
        changed = [k for k in range(len(wb)) if wa[k] != wb[k]]
        m = np.zeros((max_len, max_len), dtype=np.float64)
        if changed:
            sa, sb = _word_token_spans(base, tokenizer), _word_token_spans(other, tokenizer)
            src = [np.array(sa[k]) for k in changed]
            dst = [np.array(sb[k]) for k in changed]
        else:
            src, dst = [], []
        i = j = nxt = 0
        while i < max_len and j < max_len:
            if nxt < len(src) and src[nxt][0] == i:
                s, t = src[nxt], dst[nxt]
                if len(s) == len(t):
                    m[s, t] = 1.0
                else:
                    for tt in t:
                        m[s, tt] = 1.0 / len(t)
                i, j, nxt = i + len(s), j + len(t), nxt + 1
            elif nxt < len(src):
                m[i, j] = 1.0
                i, j = i + 1, j + 1
            else:
                m[j, j] = 1.0
                i, j = i + 1, j + 1
        out.append(torch.from_numpy(m).float())
    return torch.stack(out)


def time_words_alpha(prompts: Sequence[str], num_steps: int,
                     cross_replace_steps: Union[float, Tuple[float, float], Dict[str, Tuple[float, float]]],
                     tokenizer, max_num_words: int = MAX_WORDS) -> torch.Tensor:
    """(num_steps+1, len(prompts)-1, 1, 1, 77) per-step, per-token blend weights (p2p_utils.py:23-33,55-73)."""
    spec = cross_replace_steps if isinstance(cross_replace_steps, dict) else {"default_": cross_replace_steps}
    spec = dict(spec)
    spec.setdefault("default_", (0.0, 1.0))
    n_edit = len(prompts) - 1
    table = torch.zeros(num_steps + 1, n_edit, max_num_words)

    def window(bounds):
        lo, hi = (0, bounds) if isinstance(bounds, float) else bounds
        return int(lo * table.shape[0]), int(hi * table.shape[0])

    lo, hi = window(spec["default_"])
    table[lo:hi] = 1.0
    for word, bounds in spec.items():
        if word == "default_":
            continue
        lo, hi = window(bounds)
        for e in range(n_edit):
            text = prompts[e + 1]
            positions = [p for k, w in enumerate(text.split(" ")) if w == word
                         for p in _word_token_spans(text, tokenizer)[k]]
            if positions:
                idx = torch.tensor(positions)
                table[:, e, idx] = 0.0
                table[lo:hi, e, idx] = 1.0
    return table.reshape(num_steps + 1, n_edit, 1, 1, max_num_words)


class AttentionReplace:
    """Drop-in for src.prompt_attention.p2p_attention.AttentionReplace on the fused path."""

    def __init__(self, prompts, num_steps: int, cross_replace_steps, self_replace_steps, width, height,
                 local_blend=None, tokenizer=None, device=None, dtype=None):
        self.batch_size = len(prompts)
        if self.batch_size != 2:
            raise ValueError("the OMG pipelines run exactly two prompts (layout image, edited image)")
        self.cross_replace_alpha = time_words_alpha(prompts, num_steps, cross_replace_steps, tokenizer)
        if isinstance(self_replace_steps, float):
            self_replace_steps = 0, self_replace_steps
        self.num_self_replace = int(num_steps * self_replace_steps[0]), int(num_steps * self_replace_steps[1])
        self.mapper = replacement_mapper(prompts, tokenizer)
        self.local_blend = local_blend
        self.width, self.height = width, height
        self.num_att_layers = -1
        self.low_resource = False
        self.device, self.dtype = device, dtype
        self.reset()

    # --- reference-visible state -------------------------------------------------------------------------
    def reset(self):
        self.cur_step = 0
        self.cur_att_layer = 0

    @property
    def num_uncond_att_layers(self):
        return 0

    def between_steps(self):
        return

    def step_callback(self, x_t):
        return x_t

    def __call__(self, attn, is_cross: bool, place_in_unet: str):
        raise RuntimeError("omg_b200 never materialises attention probabilities; the fused attention kernel consumes "
                           "AttentionReplace.self_replace_active()/cross_edit() instead of a probs callback")

    # --- fused-path interface ----------------------------------------------------------------------------
    def advance(self, n_layers: int = 1):
        """Account for `n_layers` attention calls (AttentionControl.__call__ counters, p2p_attention.py:35-39)."""
        for _ in range(n_layers):
            self.cur_att_layer += 1
            if self.cur_att_layer == self.num_att_layers + self.num_uncond_att_layers:
                self.cur_att_layer = 0
                self.cur_step += 1
                self.between_steps()

    def self_replace_active(self, n_tokens: int) -> bool:
        """replace_self_attention applies (p2p_attention.py:114-118,126)."""
        lo, hi = self.num_self_replace
        return lo <= self.cur_step < hi and n_tokens <= self.width * self.height

    def self_window_active(self) -> bool:
        lo, hi = self.num_self_replace
        return lo <= self.cur_step < hi

    def cross_edit(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """Coefficients (77x77 fp32) of the two cross-attention terms at the current step:
        coef_base = M diag(alpha) (applied to V_1 and attended with image 0's probabilities) and
        coef_keep = diag(1 - alpha) (image 1's own probabilities), or None when alpha == 1 everywhere."""
        if self.cur_step >= self.cross_replace_alpha.shape[0]:
            raise IndexError(f"cur_step {self.cur_step} beyond the {self.cross_replace_alpha.shape[0]}-row alpha table"
                             " (the reference indexes it the same way, p2p_attention.py:131)")
        alpha = self.cross_replace_alpha[self.cur_step, 0, 0, 0].float()
        m = self.mapper[0].float()
        coef_base = m * alpha[None, :]
        keep = 1.0 - alpha
        coef_keep = torch.diag(keep) if bool((keep != 0).any()) else None
        return coef_base, coef_keep
