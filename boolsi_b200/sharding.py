"""Exchange of per-rank attractor tables: ONE all_gather of fixed-size integer records
(`bsi_attractor_rec`, 8 x u64) per rank, preceded in the same buffer by a header record
(count, trajectory-steps, found, without).  Replaces the MPI collection stage (reference
mpi.py:296-332, attract.py:405-455); the reduce-by-key that follows is integer addition, so the
merged table is bit-identical for every number of ranks.

Works on any `torch.distributed` backend: NCCL with device tensors in production, gloo with CPU
tensors in the tests.
"""
RECORD_WORDS = 8
DEFAULT_CAPACITY = 4096


def exchange_tables(export, header_values, world, device, capacity=DEFAULT_CAPACITY, buffers=None):
    """all_gather every rank's records.

    :param export: callable(tensor[cap, 8] int64 on `device`) -> number of records the rank HAS
        (it writes min(count, cap) of them into the tensor)
    :param header_values: up to 7 integers to ship with the records (summed by the caller)
    :param buffers: dict reused between calls to avoid reallocating
    :return: (everyone[world, cap + 1, 8], headers[world, 8] on the CPU); row 0 of each rank is its
        header: [count, *header_values]
    """
    import torch
    import torch.distributed as dist
    buffers = buffers if buffers is not None else {}
    cap = max(capacity, buffers.get('cap', 0))
    while True:
        if buffers.get('cap') != cap:
            buffers['mine'] = torch.zeros((cap + 1, RECORD_WORDS), dtype=torch.int64, device=device)
            buffers['everyone'] = torch.zeros((world, cap + 1, RECORD_WORDS), dtype=torch.int64,
                                              device=device)
            buffers['cap'] = cap
        mine, everyone = buffers['mine'], buffers['everyone']
        count = int(export(mine[1:]))
        header = [count] + [int(v) for v in header_values]
        header += [0] * (RECORD_WORDS - len(header))
        mine[0].copy_(torch.tensor(header, dtype=torch.int64))
        dist.all_gather_into_tensor(everyone.view(world * (cap + 1), RECORD_WORDS), mine)
        headers = everyone[:, 0, :].cpu()
        need = int(headers[:, 0].max())
        if need <= cap:
            return everyone, headers
        cap = 1 << need.bit_length()          # some rank overflowed: every rank retries larger
