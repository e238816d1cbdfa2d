"""VectorQuantizer(beta, legacy=False, sane_index_shape=False, remap=None) -- SURVEY.md A.1 item 7."""
import torch
import torch.nn as nn


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, vq_embed_dim, beta, remap=None, unknown_index="random", sane_index_shape=False, legacy=True):
        super().__init__()
        assert remap is None and not sane_index_shape
        self.n_e, self.vq_embed_dim, self.beta, self.legacy = n_e, vq_embed_dim, beta, legacy
        self.embedding = nn.Embedding(n_e, vq_embed_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)

    def forward(self, z):
        z = z.permute(0, 2, 3, 1).contiguous()
        zf = z.view(-1, self.vq_embed_dim)
        idx = torch.argmin(torch.cdist(zf, self.embedding.weight), dim=1)
        z_q = self.embedding(idx).view(z.shape)
        if not self.legacy:
            loss = self.beta * torch.mean((z_q.detach() - z) ** 2) + torch.mean((z_q - z.detach()) ** 2)
        else:
            loss = torch.mean((z_q.detach() - z) ** 2) + self.beta * torch.mean((z_q - z.detach()) ** 2)
        z_q = z + (z_q - z).detach()
        return z_q.permute(0, 3, 1, 2).contiguous(), loss, (None, None, idx)

    def get_codebook_entry(self, indices, shape):
        z_q = self.embedding(indices)
        if shape is not None:
            z_q = z_q.view(shape).permute(0, 3, 1, 2).contiguous()
        return z_q
