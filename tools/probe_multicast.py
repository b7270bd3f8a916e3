"""Probe (run under torchrun on >= 2 GPUs): does this box give NVLS multicast memory to a process group?
torch.distributed._symmetric_memory: symmetric allocation + rendezvous -> multicast_ptr (0 when unsupported)."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out = {"rank": rank, "world": world}
    try:
        import torch.distributed._symmetric_memory as symm
        out["has_multicast_support"] = bool(symm.has_multicast_support(torch.device("cuda", local).type, local)) if hasattr(symm, "has_multicast_support") else None
        t = symm.empty(1 << 20, dtype=torch.float32, device="cuda")
        h = symm.rendezvous(t, dist.group.WORLD.group_name)
        out["multicast_ptr"] = int(getattr(h, "multicast_ptr", 0) or 0)
        out["buffer_ptrs"] = [int(p) for p in h.buffer_ptrs]
        out["signal_pad_ptrs"] = len(h.signal_pad_ptrs)
    except Exception as ex:  # noqa: BLE001
        out["error"] = repr(ex)[:500]
    drv = {}
    try:
        from cuda import cuda as cu
        cu.cuInit(0)
        err, dev = cu.cuDeviceGet(local)
        err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev)
        drv["CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED"] = int(v)
        err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev)
        drv["HANDLE_TYPE_FABRIC_SUPPORTED"] = int(v)
        err, v = cu.cuDeviceGetAttribute(cu.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev)
        drv["HANDLE_TYPE_POSIX_FD_SUPPORTED"] = int(v)
    except Exception as ex:  # noqa: BLE001
        drv["error"] = repr(ex)[:300]
    out["driver"] = drv
    gathered = [None] * world
    dist.all_gather_object(gathered, out)
    if rank == 0:
        print(json.dumps(gathered, indent=1))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
