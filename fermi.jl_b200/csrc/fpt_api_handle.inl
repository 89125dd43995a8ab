// libfermi_pt_b200.so host driver, part of the one translation unit fpt_api.cu: handle life cycle: per-GPU state, NCCL cliques, staging-thread pool, tunables.

// ---- handle life cycle -----------------------------------------------------------------------------------------------------
// every device buffer of one GPU of a handle
static std::vector<DevBuf*> all_bufs(Dev* d)
{
    std::vector<DevBuf*> v = {&d->Pt, &d->Qt, &d->OV2, &d->T1d, &d->fo, &d->fv, &d->partials, &d->counter, &d->out, &d->prof, &d->blocktab,
                              &d->sT1, &d->sT2, &d->sOOOV, &d->sOVOV, &d->sChunk[0], &d->sChunk[1], &d->sTri, &d->sTri2, &d->sBOO, &d->sBOV, &d->sBVV,
                              &d->xTau, &d->xSlab, &d->xNew, &d->xOVOV, &d->ringtab,
                              &d->sCo, &d->sCv, &d->aoDense, &d->sIdx, &d->sVals, &d->aoQ1, &d->aoQ2v, &d->aoQ2o, &d->aoQ3vv, &d->aoQ3vo,
                              &d->aoQ3oo, &d->aoOVVV, &d->aoOOOV, &d->aoOVOV, &d->aoFlag};
    for (DevBuf& b : d->sPhase) v.push_back(&b);
    for (DevBuf& b : d->f32in) v.push_back(&b);
    for (DevBuf& b : d->f32wide) v.push_back(&b);
    return v;
}

static void dev_destroy(Dev* d)
{
    if (!d) return;
    cudaSetDevice(d->dev);
    if (d->comm) nccl_api().CommDestroy(d->comm);
    const std::vector<DevBuf*> bufs = all_bufs(d);
    for (DevBuf* b : bufs) b->release();
    cudaEvent_t evs[] = {d->ev0[0], d->ev1[0], d->ev0[1], d->ev1[1], d->ev0[2], d->ev1[2], d->ev0[3], d->ev1[3], d->ev_copy, d->ev_start, d->ev_free[0], d->ev_free[1]};
    for (cudaEvent_t e : evs)
        if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : d->tl)
        if (e) cudaEventDestroy(e);
    if (d->stream) cudaStreamDestroy(d->stream);
    if (d->copy) cudaStreamDestroy(d->copy);
    delete d;
}

static int dev_init(Dev* d)
{
    cudaDeviceProp prop;
    CK(cudaSetDevice(d->dev));
    CK(cudaGetDeviceProperties(&prop, d->dev));
    if (prop.major < 10)
        return fail("fpt_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", d->dev, prop.major, prop.minor);
    d->n_sm = prop.multiProcessorCount;
    d->total_mem = prop.totalGlobalMem;
    CK(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&d->copy, cudaStreamNonBlocking));
    static_assert(MAX_PHASES == 4, "dev_destroy lists the phase events and buffers one by one");
    for (int t = 0; t < MAX_PHASES; t++) {
        CK(cudaEventCreate(&d->ev0[t]));
        CK(cudaEventCreate(&d->ev1[t]));
    }
    CK(cudaEventCreateWithFlags(&d->ev_copy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_free[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&d->ev_free[1], cudaEventDisableTiming));
    for (int t = 0; t < NTL; t++) CK(cudaEventCreate(&d->tl[t]));
    CK(cudaFuncSetAttribute(triples_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES_SMEM_BYTES));
#ifdef FPT_WITH_VARIANT2
    CK(cudaFuncSetAttribute(triples_kernel2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
    CK(cudaFuncSetAttribute(triples_kernel2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TRIPLES2_SMEM_BYTES));
#endif
    CK(gemm_tn_set_attributes<EPI_COLMAJOR>());
    CK(gemm_tn_set_attributes<EPI_PT>());
    CK(gemm_tn_set_attributes<EPI_QT_HOLE>());
    CK(gemm_tn_set_attributes<EPI_OV2>());
    CK(gemm_tn_set_attributes<EPI_LADDER_SLAB>());
    CK(gemm_tn_set_attributes<EPI_LADDER_OUT>());
    return 0;
}

static int dev_create(int dev, int idx, Dev** out)
{
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail("fpt_create: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
    if (dev < 0) CK(cudaGetDevice(&dev));
    if (dev >= ndev) return fail("fpt_create: device %d out of range (have %d)", dev, ndev);
    Dev* d = new Dev();
    d->dev = dev;
    d->idx = idx;
    if (dev_init(d)) { dev_destroy(d); return 1; }
    *out = d;
    return 0;
}

static int default_host_threads(int share)
{
    if (const char* s = getenv("FERMI_PT_B200_THREADS")) {
        const int n = atoi(s);
        if (n >= 1) return n > 64 ? 64 : n;
    }
    int n = 1;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = CPU_COUNT(&set);
    n /= (share > 0 ? share : 1);
    return n < 1 ? 1 : (n > 16 ? 16 : n);
}

// pool of staging threads, pinned bounce ring, pinned result word
static int handle_finish(fpt_handle* h)
{
    CK(cudaSetDevice(h->devs[0]->dev));
    if (const char* nt = getenv("FERMI_PT_B200_NT")) h->pool.nt_stores = atoi(nt) != 0;
    if (const char* kb = getenv("FERMI_PT_B200_PIECE_KB")) {
        const long n = atol(kb);
        if (n >= 64 && n <= 65536) h->pool.PIECE = (size_t)n << 10;
    }
    CK(h->pool.start(default_host_threads(h->rank_mode ? h->world : 1), (int)h->devs.size()));
    CK(cudaHostAlloc((void**)&h->res_pinned, OUT_DOUBLES * sizeof(double), cudaHostAllocPortable));
    return 0;
}

extern "C" int fpt_destroy(fpt_handle* h)
{
    if (!h) return 0;
    DeviceGuard guard;
    for (Dev* d : h->devs) {
        cudaSetDevice(d->dev);
        cudaDeviceSynchronize();
    }
    h->pool.stop();
    if (h->res_pinned) cudaFreeHost(h->res_pinned);
    for (Dev* d : h->devs) dev_destroy(d);
    delete h;
    return 0;
}

extern "C" int fpt_create(int ngpu, const int* devices, fpt_handle** out)
{
    if (!out) return fail("fpt_create: out is NULL");
    *out = nullptr;
    if (ngpu < 1 || ngpu > MAX_WORLD) return fail("fpt_create: ngpu=%d out of range", ngpu);
    if (ngpu > 1 && !devices) return fail("fpt_create: a device list is required for ngpu > 1");
    DeviceGuard guard;
    fpt_handle* h = new fpt_handle();
    for (int k = 0; k < ngpu; k++) {
        Dev* d = nullptr;
        if (dev_create(devices ? devices[k] : -1, k, &d)) { fpt_destroy(h); return 1; }
        d->grank = k;
        h->devs.push_back(d);
    }
    h->world = ngpu;
    if (ngpu > 1) {
        // single-process multi-GPU: one NCCL clique over NVLink
        if (nccl_load()) { fpt_destroy(h); return 1; }
        std::vector<ncclComm_t> comms(ngpu);
        std::vector<int> devs(ngpu);
        for (int k = 0; k < ngpu; k++) devs[k] = h->devs[k]->dev;
        ncclResult_t r = nccl_api().CommInitAll(comms.data(), ngpu, devs.data());
        if (r != ncclSuccess) {
            fail("ncclCommInitAll failed: %s", nccl_api().GetErrorString(r));
            fpt_destroy(h);
            return 1;
        }
        for (int k = 0; k < ngpu; k++) h->devs[k]->comm = comms[k];
    }
    if (handle_finish(h)) { fpt_destroy(h); return 1; }
    *out = h;
    return 0;
}

extern "C" int fpt_nccl_unique_id(void* id128)
{
    if (!id128) return fail("fpt_nccl_unique_id: NULL argument");
    if (nccl_load()) return 1;
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI carries the NCCL id as 128 bytes");
    ncclUniqueId id;
    NCK(nccl_api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return 0;
}

extern "C" int fpt_create_rank(int device, int rank, int world, const void* id128, fpt_handle** out)
{
    if (!out) return fail("fpt_create_rank: out is NULL");
    *out = nullptr;
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) return fail("fpt_create_rank: invalid rank %d of %d (at most %d GPUs)", rank, world, MAX_WORLD);
    if (world > 1 && !id128) return fail("fpt_create_rank: the NCCL id is required for world > 1");
    DeviceGuard guard;
    fpt_handle* h = new fpt_handle();
    Dev* d = nullptr;
    if (dev_create(device, 0, &d)) { fpt_destroy(h); return 1; }
    d->grank = rank;
    h->devs.push_back(d);
    h->world = world;
    h->rank_mode = true;
    if (world > 1) {
        if (nccl_load()) { fpt_destroy(h); return 1; }
        ncclUniqueId id;
        memcpy(&id, id128, sizeof id);
        cudaSetDevice(d->dev);
        ncclResult_t r = nccl_api().CommInitRank(&d->comm, world, id, rank);
        if (r != ncclSuccess) {
            d->comm = nullptr;
            fail("ncclCommInitRank failed: %s", nccl_api().GetErrorString(r));
            fpt_destroy(h);
            return 1;
        }
    }
    if (handle_finish(h)) { fpt_destroy(h); return 1; }
    *out = h;
    return 0;
}

extern "C" int fpt_set_symmetric_inputs(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_symmetric_inputs: NULL handle");
    h->sym_inputs = on ? 1 : 0;
    return 0;
}

extern "C" int fpt_set_adaptive_shards(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_adaptive_shards: NULL handle");
    h->adaptive = on ? 1 : 0;
    for (ShardCal& c : h->cal) c = ShardCal{};
    return 0;
}

extern "C" int fpt_set_deterministic(fpt_handle* h, int on)
{
    if (!h) return fail("fpt_set_deterministic: NULL handle");
    h->deterministic = on ? 1 : 0;
    return 0;
}

extern "C" int fpt_set_df_ring(fpt_handle* h, int block)
{
    if (!h) return fail("fpt_set_df_ring: NULL handle");
    if (block < -1 || block > 64) return fail("fpt_set_df_ring: block=%d out of range (-1 .. 64)", block);
    h->df_ring = block;
    return 0;
}

extern "C" int fpt_device_bytes(fpt_handle* h, double* bytes)
{
    if (!h || !bytes) return fail("fpt_device_bytes: NULL argument");
    const std::vector<DevBuf*> bufs = all_bufs(h->devs[0]);
    double n = 0.0;
    for (DevBuf* b : bufs) n += (double)b->cap;
    *bytes = n;
    return 0;
}

extern "C" int fpt_set_host_threads(fpt_handle* h, int n)
{
    if (!h) return fail("fpt_set_host_threads: NULL handle");
    if (n < 1 || n > 64) return fail("fpt_set_host_threads: n=%d out of range (1..64)", n);
    DeviceGuard guard;
    for (Dev* d : h->devs) {   // no DMA may still be reading the slots that are about to be freed
        CK(cudaSetDevice(d->dev));
        CK(cudaStreamSynchronize(d->copy));
    }
    CK(h->pool.start(n, (int)h->devs.size()));
    return 0;
}
