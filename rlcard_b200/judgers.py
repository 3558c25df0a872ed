"""The judgers / encoders of the hot path as standalone batched operators (device tensors in, device tensors out).

Mirrors of the reference's stateless helpers: ``compare_hands`` (games/limitholdem/utils.py:526-569),
``LeducholdemJudger.judge_game`` (games/leducholdem/judger.py:12-64), ``DoudizhuJudger.playable_cards_from_hand``
(games/doudizhu/judger.py:124-258) / ``get_gt_cards`` (games/doudizhu/utils.py:225-262) and UNO's ``encode_hand`` /
``encode_target`` (games/uno/utils.py:86-127).  They run the same device functions the env kernels use.
"""
import ctypes as C

import torch

from ._lib import GAME_IDS, check, game_info, lib


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _dev(t, dtype):
    if t.device.type != 'cuda':
        raise ValueError('judger operators take CUDA tensors (there is no CPU fallback)')
    return t.to(dtype).contiguous()


def compare_hands(cards):
    """cards uint8 [n, P, 7] (13*suit + rank ids of card2index.json; first card 255 = folded) -> winners uint8 [n, P]."""
    cards = _dev(cards, torch.uint8)
    n, P, seven = cards.shape
    assert seven == 7
    out = torch.empty((n, P), dtype=torch.uint8, device=cards.device)
    with torch.cuda.device(cards.device):
        check(lib().rlc_judge_holdem(_p(cards), n, P, _p(out), _stream(cards.device)))
    return out


def judge_leduc(cases):
    """cases int32 [n, 7] = (rank0, rank1, public rank or -1, chips0, chips1, folded0, folded1) -> payoffs float32 [n, 2]."""
    cases = _dev(cases, torch.int32)
    out = torch.empty((cases.shape[0], 2), dtype=torch.float32, device=cases.device)
    with torch.cuda.device(cases.device):
        check(lib().rlc_judge_leduc(_p(cases), cases.shape[0], _p(out), _stream(cases.device)))
    return out


def doudizhu_playable(hands, targets=None):
    """hands uint8 [n, 15] rank counts, targets int32 [n] (action id to beat, < 0 = lead) -> bit-packed legal sets
    int32 [n, mask_words = 860] (bit a % 32 of word a // 32)."""
    from .vec_env import _upload_doudizhu_tables
    hands = _dev(hands, torch.uint8)
    _upload_doudizhu_tables(lib(), hands.device)
    if targets is not None:
        targets = _dev(targets, torch.int32)
    out = torch.empty((hands.shape[0], game_info('doudizhu').mask_words), dtype=torch.int32, device=hands.device)
    with torch.cuda.device(hands.device):
        check(lib().rlc_judge_doudizhu(_p(hands), _p(targets), hands.shape[0], _p(out), _stream(hands.device)))
    return out


def uno_encode(hands, targets):
    """hands uint8 [n, 32] card codes 15*colour + trait (255 = empty), targets uint8 [n] -> obs uint8 [n, 4, 4, 15]."""
    hands, targets = _dev(hands, torch.uint8), _dev(targets, torch.uint8)
    assert hands.shape[1] == 32
    out = torch.empty((hands.shape[0], 240), dtype=torch.uint8, device=hands.device)
    with torch.cuda.device(hands.device):
        check(lib().rlc_encode_uno(_p(hands), _p(targets), hands.shape[0], _p(out), _stream(hands.device)))
    return out.view(-1, 4, 4, 15)


__all__ = ['compare_hands', 'judge_leduc', 'doudizhu_playable', 'uno_encode', 'GAME_IDS']
