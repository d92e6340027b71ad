"""The reference's models assembled from the cti_b200 drop-ins, for end-to-end tests and the full-model bench leg:
``MCModel`` = the multiple-choice CTI model (``TanModel`` + ``build_cti``, reference src/MC/base_model.py:112-152,186-208),
``BanStudent`` = the free-form BAN student without counter (``BanModel`` + ``build_ban``, src/FFOE/base_model.py:21-66,145-163).  Host glue, not product: with
the reference tree present one calls ``cti_b200.install()`` and the reference's own builder instead.
Attribute names equal the reference's, so a reference ``state_dict`` loads unchanged."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cti_b200  # noqa: E402


class WordEmbedding(nn.Module):
    """reference src/language_model.py:11-47 (op 'c': a trainable and a frozen table, concatenated)."""

    def __init__(self, ntoken, emb_dim, dropout, op=''):
        super().__init__()
        self.op = op
        self.emb = nn.Embedding(ntoken + 1, emb_dim, padding_idx=ntoken)
        if 'c' in op:
            self.emb_ = nn.Embedding(ntoken + 1, emb_dim, padding_idx=ntoken)
            self.emb_.weight.requires_grad = False
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        emb = self.emb(x)
        if 'c' in self.op:
            emb = torch.cat((emb, self.emb_(x)), 2)
        return self.dropout(emb)


class MCModel(nn.Module):
    def __init__(self, ntoken, v_dim, num_hid, h_mm, rank, gamma, op='c', activation='relu', dropout=0.5):
        super().__init__()
        args = type("Args", (), {"activation": activation, "dropout": dropout})()
        din = 300 if 'c' not in op else 600
        self.glimpse = gamma
        self.w_emb = WordEmbedding(ntoken, 300, .0, op)
        self.q_emb = cti_b200.QuestionEmbedding(din, num_hid, 1, False, .0)
        self.wa_emb = WordEmbedding(ntoken, 300, .0, op)
        self.ans_emb = cti_b200.QuestionEmbedding(din, num_hid, 1, False, .0)
        self.v_att = cti_b200.TriAttention(v_dim, num_hid, num_hid, h_mm, 1, rank, gamma, 1, dropout=[.2, .5])
        self.t_net = nn.ModuleList([cti_b200.TCNet(v_dim, num_hid, num_hid, h_mm, 1, rank, 1, dropout=[.2, .5], k=2)
                                    for _ in range(gamma)])
        self.q_prj = nn.ModuleList([cti_b200.FCNet([num_hid, num_hid], '', .2) for _ in range(gamma)])
        self.a_prj = nn.ModuleList([cti_b200.FCNet([num_hid, num_hid], '', .2) for _ in range(gamma)])
        self.classifier = cti_b200.SimpleClassifier(num_hid, num_hid * 2, 2, args)

    def forward(self, v, b, q, ans):
        """v [rows (or rows / 4), objs, dim], q [rows, 12] token ids, ans [rows, 6] token ids -> (logits [rows, 2], att)."""
        q_emb = self.q_emb.forward_all(self.w_emb(q))
        ans_emb = self.ans_emb.forward_all(self.wa_emb(ans))
        att, _ = self.v_att(v, q_emb, ans_emb)
        for g in range(self.glimpse):
            b_emb = self.t_net[g].forward_with_weights(v, q_emb, ans_emb, att[:, :, :, :, g])
            q_emb = self.q_prj[g](b_emb.unsqueeze(1)) + q_emb
            ans_emb = self.a_prj[g](b_emb.unsqueeze(1)) + ans_emb
        return self.classifier(q_emb.sum(1) + ans_emb.sum(1)), att


class CTIFreeForm(MCModel):
    """reference src/FFOE/base_model.py:96-134,177-200 (``CTIModel`` + FFOE ``build_cti``): the free-form CTI teacher --
    the MC model with ``n_ans`` classes, the attention registered as ``t_att`` and ``forward(v, q, ans)`` -> logits."""

    def __init__(self, ntoken, v_dim, num_hid, h_mm, rank, gamma, n_ans, op='c', activation='relu', dropout=0.5):
        super().__init__(ntoken, v_dim, num_hid, h_mm, rank, gamma, op, activation, dropout)
        args = type("Args", (), {"activation": activation, "dropout": dropout})()
        self.classifier = cti_b200.SimpleClassifier(num_hid, num_hid * 2, n_ans, args)
        self.t_att = self.v_att
        del self.v_att

    def forward(self, v, q, ans):
        q_emb = self.q_emb.forward_all(self.w_emb(q))
        ans_emb = self.ans_emb.forward_all(self.wa_emb(ans))
        att, _ = self.t_att(v, q_emb, ans_emb)
        for g in range(self.glimpse):
            b_emb = self.t_net[g].forward_with_weights(v, q_emb, ans_emb, att[:, :, :, :, g])
            q_emb = self.q_prj[g](b_emb.unsqueeze(1)) + q_emb
            ans_emb = self.a_prj[g](b_emb.unsqueeze(1)) + ans_emb
        return self.classifier(q_emb.sum(1) + ans_emb.sum(1))


class BanStudent(nn.Module):
    """reference src/FFOE/base_model.py:21-66 with ``use_counter=False`` (the distillation student, README.md:49)."""

    def __init__(self, ntoken, v_dim, num_hid, gamma, n_ans, op='c', activation='relu', dropout=0.5):
        super().__init__()
        args = type("Args", (), {"activation": activation, "dropout": dropout})()
        self.glimpse = gamma
        self.w_emb = WordEmbedding(ntoken, 300, .0, op)
        self.q_emb = cti_b200.QuestionEmbedding(300 if 'c' not in op else 600, num_hid, 1, False, .0)
        self.v_att = cti_b200.BiAttention(v_dim, num_hid, num_hid, gamma)
        self.b_net = nn.ModuleList([cti_b200.BCNet(v_dim, num_hid, num_hid, None, k=1) for _ in range(gamma)])
        self.q_prj = nn.ModuleList([cti_b200.FCNet([num_hid, num_hid], '', .2) for _ in range(gamma)])
        self.classifier = cti_b200.SimpleClassifier(num_hid, num_hid * 2, n_ans, args)

    def forward(self, v, b, q, labels=None):
        q_emb = self.q_emb.forward_all(self.w_emb(q))
        att, _ = self.v_att.forward_all(v, q_emb)
        q_list = []
        for g in range(self.glimpse):
            b_emb = self.b_net[g].forward_with_weights(v, q_emb, att[:, g, :, :])
            q_emb = self.q_prj[g](b_emb.unsqueeze(1)) + q_emb
            q_list.append(q_emb)
        return self.classifier(torch.stack(q_list, 1).sum(1).sum(1)), att
