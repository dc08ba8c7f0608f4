"""``torch_geometric.data.Data`` / ``Batch`` as attribute bags with the few methods the reference calls
(``to``, ``has_isolated_nodes``, ``num_nodes``, ``Batch.from_data_list``) -- TEST INFRASTRUCTURE."""
import torch


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        for k, v in dict(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos).items():
            if v is not None:
                setattr(self, k, v)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def keys(self):
        return [k for k in self.__dict__ if not k.startswith("_")]

    def to(self, device, *args, **kwargs):
        for k in self.keys():
            v = getattr(self, k)
            if torch.is_tensor(v):
                setattr(self, k, v.to(device, *args, **kwargs))
        return self

    @property
    def num_nodes(self):
        for k in ("x", "pos", "batch"):
            v = getattr(self, k, None)
            if torch.is_tensor(v):
                return v.shape[0]
        # PyG's inference order continues with node-level attributes (names containing "node": the offline n-body graphs
        # carry node_feat / node_attr and no x / pos) and ends at the largest index of edge_index
        for k in self.keys():
            v = getattr(self, k)
            if "node" in k and torch.is_tensor(v):
                return v.shape[0]
        ei = getattr(self, "edge_index", None)
        if torch.is_tensor(ei) and ei.numel():
            return int(ei.max()) + 1
        return None

    def has_isolated_nodes(self) -> bool:
        ei = self.edge_index
        n = self.num_nodes
        seen = torch.zeros(n, dtype=torch.bool, device=ei.device)
        seen[ei.reshape(-1)] = True
        return bool((~seen).any())


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list):
        out = cls()
        keys = data_list[0].keys()
        sizes = [d.num_nodes for d in data_list]
        for k in keys:
            vals = [getattr(d, k) for d in data_list]
            if k == "edge_index":
                offs, acc = [], 0
                for n in sizes:
                    offs.append(acc)
                    acc += n
                setattr(out, k, torch.cat([v + o for v, o in zip(vals, offs)], dim=1))
            elif torch.is_tensor(vals[0]):
                setattr(out, k, torch.cat(vals, dim=0))
            else:
                setattr(out, k, vals)
        out.batch = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes)])
        out.batch = out.batch.to(getattr(out, "pos", out.batch).device)
        out.ptr = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.long)
        out.num_graphs = len(data_list)
        return out
