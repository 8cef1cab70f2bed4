"""Exchange of per-rank attractor tables: ONE all_gather of fixed-size integer records
(`bsi_attractor_rec`, 8 x u64) per rank.  Row 0 of a rank's payload is a header record written on
the device -- words [count, trajectory-steps, found, without, 0 (n), overflow flag, 0, 0] -- and
rows 1.. are its records.  Replaces the MPI collection stage (reference mpi.py:296-332,
attract.py:405-455); the reduce-by-key that follows is integer addition, so the merged table is
bit-identical for every number of ranks.

Works on any `torch.distributed` backend: NCCL with device tensors in production, gloo with CPU
tensors in the tests.
"""
RECORD_WORDS = 8
DEFAULT_CAPACITY = 1024


def exchange_tables(fill_payload, world, device, capacity=DEFAULT_CAPACITY, buffers=None):
    """all_gather every rank's payload.

    :param fill_payload: callable(tensor[cap + 1, 8] int64 on `device`): writes the header row and
        up to `cap` records (asynchronously is fine -- the collective is stream-ordered after it)
    :param buffers: dict reused between calls to avoid reallocating
    :return: (everyone[world, cap + 1, 8], headers[world, 8] on the CPU)
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
        fill_payload(mine)
        dist.all_gather_into_tensor(everyone.view(world * (cap + 1), RECORD_WORDS), mine)
        headers = everyone[:, 0, :].cpu()             # the one host synchronisation of the exchange
        need = int(headers[:, 0].max())
        if need <= cap:
            return everyone, headers
        cap = 1 << need.bit_length()          # some rank overflowed: every rank retries larger
