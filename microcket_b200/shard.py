"""Multi-GPU plumbing of the dedup/binning exchange.

The data path has ONE exchange step.  On GPUs it is the library's own kernel over NVLink peer memory
(mk_xchg_*, csrc/xchg.cu): torch.distributed only carries the 128-byte cudaIpc handles at start-up (connect_xchg).
The collective form below (owner partition on the device, counts exchanged, one all_to_all_single) is kept for boxes
without peer access (MICROCKET_XCHG=nccl) and is what the CPU tests drive with gloo.  Parsing needs no collective
(shards are cut at read-group boundaries) and the COO counts are owner-computed, so nothing else crosses NVLink.
"""


def connect_xchg(torch, dist, xchg, device):
    """Bootstrap of the NVLink peer-memory exchange (csrc/xchg.cu): all-gather the ranks' 128-byte cudaIpc handles."""
    mine = torch.frombuffer(bytearray(xchg.handle()), dtype=torch.uint8).to(device)
    allh = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allh, mine)
    xchg.connect(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))


def exchange_counts(torch, dist, counts, device):
    """counts[r] = elements this rank sends to rank r → list of elements received from every rank."""
    send = torch.tensor(counts, dtype=torch.int64, device=device)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    return [int(x) for x in recv.tolist()]


def exchange_segments(torch, dist, send_buf, send_counts, recv_buf, recv_counts, elem_bytes=16):
    """send_buf: uint8 tensor with the segments in rank order; recv_buf gets the received segments in rank order."""
    n_send, n_recv = sum(send_counts), sum(recv_counts)
    a = send_buf[:n_send * elem_bytes].view(n_send, elem_bytes)
    b = recv_buf[:n_recv * elem_bytes].view(n_recv, elem_bytes)
    dist.all_to_all_single(b, a, output_split_sizes=recv_counts, input_split_sizes=send_counts)
    return n_recv


def exchange_pairs(mk, torch, dist, ws, pairs, n, recv, cap_pairs, res, stream):
    """Partition `n` packed pairs (uint8 cuda tensor) by owner and all-to-all them.  → (n_received, tensor holding them)"""
    world = dist.get_world_size()
    part = recv                                  # partition into `recv`, receive back into `pairs`' storage
    counts = ws.partition(pairs.data_ptr(), n, world, res, part.data_ptr(), stream=stream)
    rc = exchange_counts(torch, dist, counts, pairs.device)
    if sum(rc) > cap_pairs or sum(rc) * 16 > pairs.numel():
        raise RuntimeError(f"rank receives {sum(rc)} pairs, capacity {cap_pairs}")
    n_recv = exchange_segments(torch, dist, part, counts, pairs, rc)
    return n_recv, pairs
