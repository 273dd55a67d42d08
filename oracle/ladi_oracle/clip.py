"""TEST INFRASTRUCTURE ONLY (see oracle/ladi_oracle/__init__.py) -- fp32 CPU restatement of the CLIP towers that feed the try-on
path:

  * `ClipTextEncoder`  plays `transformers==4.27.3` `CLIPTextModel` with exactly the attribute surface that the reference's
    /root/reference/src/utils/encode_text_word_embedding.py:6-72 touches (`text_model.embeddings.{position_ids, token_embedding,
    position_embedding}`, `text_model._build_causal_attention_mask`, `text_model.encoder(inputs_embeds=..., causal_attention_mask=...)`,
    `text_model.final_layer_norm`), so that file runs UNMODIFIED on it (the installed transformers 5.5 no longer has
    `_build_causal_attention_mask`, so the reference function cannot run on the library class).
  * `ClipVisionEncoder` plays `CLIPVisionModelWithProjection(...)(pixel_values).last_hidden_state` (src/inference.py:269-273).
  * `encode_text_word_embedding` restates the reference function (same argument meaning) for use without /root/reference.

Parity status: the layer arithmetic is pinned against the installed `transformers` 5.5 `CLIPTextModel` / `CLIPVisionModel` with shared
random weights (tests/test_oracle_pins.py: identical state-dict keys, outputs equal to fp32 rounding); the '$' substitution is pinned
against the reference function itself run on `ClipTextEncoder` (tests/golden/make_golden.py -> tests/golden/clip_text_small.npz).
Architecture constants (SD-2 text encoder: 23 layers, 1024 wide, 16 heads, MLP 4096, GELU(erf); ViT-H/14: 32 layers, 1280 wide, 16
heads of 80, MLP 5120, 224 px / patch 14) are the public model configs named at src/inference.py:123-138 and hubconf.py:17 -- recalled,
not fetched (no network).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

DOLLAR_ID = 259  # '$' in the CLIP vocabulary (encode_text_word_embedding.py:13)


class _Attn(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            setattr(self, n, nn.Linear(dim, dim))


class ClipLayer(nn.Module):
    """transformers 4.27.3 CLIPEncoderLayer written out: pre-LN, q scaled by d^-0.5 before QK^T, additive masks, fp32 softmax."""

    def __init__(self, dim, heads, mlp, eps=1e-5):
        super().__init__()
        self.self_attn = _Attn(dim, heads)
        self.layer_norm1 = nn.LayerNorm(dim, eps=eps)
        self.mlp = nn.Module()
        self.mlp.fc1 = nn.Linear(dim, mlp)
        self.mlp.fc2 = nn.Linear(mlp, dim)
        self.layer_norm2 = nn.LayerNorm(dim, eps=eps)

    def forward(self, x, attention_mask=None, causal_attention_mask=None):
        b, n, c = x.shape
        a = self.self_attn
        hd = c // a.heads
        y = self.layer_norm1(x)
        sp = lambda t: t.view(b, n, a.heads, hd).transpose(1, 2)
        s = (sp(a.q_proj(y)) * hd ** -0.5) @ sp(a.k_proj(y)).transpose(-1, -2)
        if causal_attention_mask is not None:
            s = s + causal_attention_mask
        if attention_mask is not None:
            s = s + attention_mask
        o = torch.softmax(s, dim=-1) @ sp(a.v_proj(y))
        x = x + a.out_proj(o.transpose(1, 2).reshape(b, n, c))
        return x + self.mlp.fc2(F.gelu(self.mlp.fc1(self.layer_norm2(x))))


class _EncoderOutput(tuple):
    """(last_hidden_state,) with the attribute names BaseModelOutput carries (encode_text_word_embedding.py:55,68-69)."""

    def __new__(cls, h):
        o = super().__new__(cls, (h,))
        o.last_hidden_state, o.hidden_states, o.attentions = h, None, None
        return o


class _Encoder(nn.Module):
    def __init__(self, dim, heads, mlp, layers):
        super().__init__()
        self.layers = nn.ModuleList([ClipLayer(dim, heads, mlp) for _ in range(layers)])

    def forward(self, inputs_embeds, attention_mask=None, causal_attention_mask=None, output_attentions=None,
                output_hidden_states=None, return_dict=None):
        h = inputs_embeds
        for layer in self.layers:
            h = layer(h, attention_mask, causal_attention_mask)
        return _EncoderOutput(h)


class _Output:
    def __init__(self, last_hidden_state, pooler_output):
        self.last_hidden_state, self.pooler_output = last_hidden_state, pooler_output

    def __getitem__(self, i):
        return (self.last_hidden_state, self.pooler_output)[i]


class _TextEmbeddings(nn.Module):
    def __init__(self, vocab, dim, max_pos):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab, dim)
        self.position_embedding = nn.Embedding(max_pos, dim)
        self.register_buffer("position_ids", torch.arange(max_pos).unsqueeze(0), persistent=False)


class _TextTransformer(nn.Module):
    def __init__(self, vocab, dim, heads, layers, mlp, max_pos):
        super().__init__()
        self.embeddings = _TextEmbeddings(vocab, dim, max_pos)
        self.encoder = _Encoder(dim, heads, mlp, layers)
        self.final_layer_norm = nn.LayerNorm(dim, eps=1e-5)

    @staticmethod
    def _build_causal_attention_mask(bsz, seq_len, dtype):
        m = torch.empty(bsz, seq_len, seq_len, dtype=dtype)
        m.fill_(torch.finfo(dtype).min)
        m.triu_(1)
        return m.unsqueeze(1)

    def forward(self, input_ids):
        b, n = input_ids.shape
        e = self.embeddings
        h = e.token_embedding(input_ids) + e.position_embedding(e.position_ids[:, :n])
        h = self.encoder(inputs_embeds=h, causal_attention_mask=self._build_causal_attention_mask(b, n, h.dtype))[0]
        h = self.final_layer_norm(h)
        return _Output(h, h[torch.arange(b), input_ids.to(torch.int).argmax(dim=-1)])


class ClipTextEncoder(nn.Module):
    def __init__(self, vocab=49408, dim=1024, heads=16, layers=23, mlp=4096, max_pos=77):
        super().__init__()
        self.text_model = _TextTransformer(vocab, dim, heads, layers, mlp, max_pos)

    @property
    def dtype(self):
        return self.text_model.final_layer_norm.weight.dtype

    def forward(self, input_ids, attention_mask=None):
        return self.text_model(input_ids)


def encode_text_word_embedding(text_encoder, input_ids, word_embeddings, num_vstar=1):
    """Restatement of /root/reference/src/utils/encode_text_word_embedding.py:6-72: in every prompt that contains '$' (id 259), the
    `num_vstar` token embeddings starting at its FIRST '$' are replaced by that prompt's pseudo-word embeddings; then the causal
    CLIP text transformer + final LayerNorm; pooled = the row at argmax(input_ids) (the EOT token)."""
    tm = text_encoder.text_model
    b, n = input_ids.shape
    h = tm.embeddings.token_embedding(input_ids)
    if word_embeddings is not None:
        we = word_embeddings.to(h.dtype)
        if we.dim() == 2:
            we = we.unsqueeze(1)
        for i in range(b):
            hit = (input_ids[i] == DOLLAR_ID).nonzero()
            if len(hit) > 0:
                f = int(hit[0])
                h[i, f:f + num_vstar] = we[i, :num_vstar]
    h = h + tm.embeddings.position_embedding(tm.embeddings.position_ids[:, :n])
    h = tm.encoder(inputs_embeds=h, causal_attention_mask=tm._build_causal_attention_mask(b, n, h.dtype))[0]
    h = tm.final_layer_norm(h)
    return _Output(h, h[torch.arange(b), input_ids.to(torch.int).argmax(dim=-1)])


class _VisionEmbeddings(nn.Module):
    def __init__(self, dim, image, patch):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(dim))
        self.patch_embedding = nn.Conv2d(3, dim, kernel_size=patch, stride=patch, bias=False)
        self.position_embedding = nn.Embedding((image // patch) ** 2 + 1, dim)

    def forward(self, px):
        p = self.patch_embedding(px).flatten(2).transpose(1, 2)
        x = torch.cat([self.class_embedding.expand(px.shape[0], 1, -1), p], dim=1)
        return x + self.position_embedding.weight[None]


class _VisionTransformer(nn.Module):
    def __init__(self, dim, heads, layers, mlp, image, patch):
        super().__init__()
        self.embeddings = _VisionEmbeddings(dim, image, patch)
        self.pre_layrnorm = nn.LayerNorm(dim, eps=1e-5)  # (sic) the upstream attribute name
        self.encoder = _Encoder(dim, heads, mlp, layers)
        self.post_layernorm = nn.LayerNorm(dim, eps=1e-5)

    def forward(self, px):
        h = self.encoder(inputs_embeds=self.pre_layrnorm(self.embeddings(px)))[0]
        return _Output(h, self.post_layernorm(h[:, 0]))  # last_hidden_state is taken BEFORE post_layernorm


class ClipVisionEncoder(nn.Module):
    def __init__(self, dim=1280, heads=16, layers=32, mlp=5120, image=224, patch=14):
        super().__init__()
        self.vision_model = _VisionTransformer(dim, heads, layers, mlp, image, patch)

    def forward(self, pixel_values):
        return self.vision_model(pixel_values)
