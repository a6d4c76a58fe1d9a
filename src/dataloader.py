"""Parquet loader with the reference's interface (src/dataloader.py): ParquetDataset, BatchCollator,
MMCTRDataLoader(feature_map, data_path, item_info_path, batch_size, shuffle, num_workers, max_len).

Same contract towards the model: every parquet column is stacked into ONE float64 matrix (so scalar columns
reach the model as float64), `item_seq` is cropped to the last `max_len` ids and cast to int64,
`item_emb_d128` (B,128) float32 is looked up by `item_id`, the label is popped and returned as float32.
Difference: the multimodal vectors come from a dense matrix indexed by item_id instead of a per-batch pandas
`.loc` over object cells (the reference builds that matrix and never uses it).
"""
import numpy as np
import pandas as pd
import torch
from torch.utils.data import DataLoader, Dataset


class ParquetDataset(Dataset):
    def __init__(self, data_path):
        self.column_index = {}
        self.darray = self.load_data(data_path)

    def __len__(self):
        return self.darray.shape[0]

    def __getitem__(self, index):
        return self.darray[index, :]

    def __getitems__(self, indices):
        """Batched fetch (torch's DataLoader calls it with the whole index list of a batch): ONE fancy-index copy instead of
        batch_size row views + np.stack -- the per-row path is what bounds the reference's loader."""
        idx = np.asarray(indices, dtype=np.int64)
        if idx.size and idx[-1] - idx[0] + 1 == idx.size and (idx.size == 1 or (np.diff(idx) == 1).all()):
            return self.darray[idx[0]:idx[-1] + 1, :]          # sequential batch (shuffle=False): a view, no copy at all
        return self.darray[idx, :]

    def load_data(self, data_path):
        frame = pd.read_parquet(data_path)
        blocks, cursor = [], 0
        for name in frame.columns:
            col = frame[name]
            if col.dtype == "object":                      # list-valued column -> (N, len) block
                block = np.array(col.to_list())
                if block.ndim == 1:
                    block = block.reshape(-1, 1)
                width = block.shape[1]
                self.column_index[name] = list(range(cursor, cursor + width))
            else:
                block = col.to_numpy().reshape(-1, 1)
                width = 1
                self.column_index[name] = cursor
            cursor += width
            blocks.append(block)
        return np.column_stack(blocks)


def load_item_embedding_matrix(item_info_path):
    """(max_item_id+1, 128) float32 matrix of item_emb_d128 indexed by item_id (missing ids -> zeros)."""
    info = pd.read_parquet(item_info_path)
    ids = info["item_id"].to_numpy().astype(np.int64)
    vecs = np.stack([np.asarray(v, dtype=np.float32) for v in info["item_emb_d128"].to_list()])
    table = np.zeros((int(ids.max()) + 1, vecs.shape[1]), dtype=np.float32)
    known = np.zeros(table.shape[0], dtype=bool)
    table[ids] = vecs
    known[ids] = True
    return table, known


class BatchCollator:
    def __init__(self, feature_map, max_len, column_index, item_info_path, strict=True, with_mm=True):
        # with_mm=False: the (B,128) item_emb_d128 block is NOT materialised per batch -- the model gathers it on the GPU from the
        # resident matrix (model.attach_mm_table(collator.item_embedding_matrix)); unknown ids are still checked when strict
        self.with_mm = with_mm
        self.feature_map = feature_map
        self.max_len = max_len
        self.column_index = column_index
        self.strict = strict        # training raises on unknown item ids (reference: KeyError), inference fills zeros
        self.item_embedding_matrix, self._known = load_item_embedding_matrix(item_info_path)
        # list-valued columns occupy contiguous matrix columns: slice them (a view) instead of fancy-indexing a 100-wide copy
        self._index = {}
        for name, idx in column_index.items():
            if isinstance(idx, list) and idx == list(range(idx[0], idx[0] + len(idx))):
                idx = slice(idx[0], idx[0] + len(idx))
            self._index[name] = idx

    def lookup(self, item_ids):
        ids = np.asarray(item_ids).astype(np.int64)
        inside = (ids >= 0) & (ids < self._known.shape[0])
        ok = inside.copy()
        ok[inside] = self._known[ids[inside]]
        if self.strict and not ok.all():
            raise KeyError(f"item ids missing from item_info: {ids[~ok][:10].tolist()}")
        if ok.all():
            return self.item_embedding_matrix[ids]
        out = np.zeros((ids.shape[0], self.item_embedding_matrix.shape[1]), dtype=np.float32)
        out[ok] = self.item_embedding_matrix[ids[ok]]
        return out

    def check_known(self, item_ids):
        ids = np.asarray(item_ids).astype(np.int64)
        inside = (ids >= 0) & (ids < self._known.shape[0])
        if not inside.all() or not self._known[ids].all():
            bad = ids[~inside] if not inside.all() else ids[~self._known[ids]]
            raise KeyError(f"item ids missing from item_info: {bad[:10].tolist()}")

    def __call__(self, rows):
        # a (B, columns) block from ParquetDataset.__getitems__, or the reference's list of row vectors
        mat = torch.from_numpy(rows if isinstance(rows, np.ndarray) else np.stack(rows))
        batch = {}
        for name, idx in self._index.items():
            batch[name] = mat[:, idx]
        if self.with_mm:
            batch["item_emb_d128"] = torch.from_numpy(self.lookup(batch["item_id"].numpy()))
        elif self.strict:
            self.check_known(batch["item_id"].numpy())
        if "item_seq" in batch:
            seq = batch["item_seq"]
            if seq.shape[1] > self.max_len:
                seq = seq[:, -self.max_len:]
            batch["item_seq"] = seq.long()
        if "label" in batch:
            labels = batch.pop("label").float()
            return batch, labels
        return batch


class MMCTRDataLoader(DataLoader):
    def __init__(self, feature_map, data_path, item_info_path, batch_size=32, shuffle=False, num_workers=1, max_len=100,
                 with_mm=True, **kwargs):
        if not data_path.endswith(".parquet"):
            data_path += ".parquet"
        self.dataset = ParquetDataset(data_path)
        self.column_index = self.dataset.column_index
        collator = BatchCollator(feature_map, max_len, self.column_index, item_info_path, with_mm=with_mm)
        self.collator = collator
        super().__init__(dataset=self.dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                         collate_fn=collator, **kwargs)
