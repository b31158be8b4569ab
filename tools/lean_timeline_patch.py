"""Development aid (how the tail problems of k_weight_lean were found): patches a %globaltimer mark per warp and chunk into the
lean weight kernel and a dump of the last launch into eslam_gpu_run_steps.  Apply on a clean tree, build, run
`ESLAM_LEAN_TRACE=1 python bench.py --steps 8 --warmup 5 --no-cpu-baseline` on the GPU (writes gpurun_out/lean_trace.bin), read it with
tools/lean_timeline_report.py, then `git checkout slam-eslam_b200/csrc`.  The anchors below are exact source lines: if one no longer matches,
the replace is a no-op and the build fails or the trace stays empty."""
p='/root/repo/slam-eslam_b200/csrc/eslam_kernels.cuh'
s=open(p).read()
s=s.replace("struct FastParams { WeightCommon c; FastContacts fc; };","struct FastParams { WeightCommon c; FastContacts fc; unsigned long long* trace; };",1)
s=s.replace("""    int bin = cta_bin[blockIdx.x];
    uint32_t phase = 0;
    while (true) {""","""    int bin = cta_bin[blockIdx.x];
    uint32_t phase = 0;
    int tslot = 0;
    auto tmark = [&]() { if (P.trace && lane == 0 && tslot < 16) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); P.trace[((size_t)blockIdx.x * 32 + (tid >> 5)) * 16 + tslot++] = t; } };
    tmark();
    while (true) {""",1)
s=s.replace("""                if (have) finish_particle(C, cur, i, ev, z, zs2, C.st.f[cur][F_W][i]);
            }""","""                if (have) finish_particle(C, cur, i, ev, z, zs2, C.st.f[cur][F_W][i]);
                tmark();
            }""",1)
s=s.replace("""        bin = s_pick[1];
        if (bin < 0) break;
        __syncthreads();
    }
}""","""        bin = s_pick[1];
        tmark();
        if (bin < 0) break;
        __syncthreads();
    }
}""",1)
open(p,'w').write(s)
p='/root/repo/slam-eslam_b200/csrc/eslam_gpu.cu'
s=open(p).read()
s=s.replace("    fc.c0x = 402653184.0 - P.map.ox * P.map.inv_sx;                       // 1.5 * 2^28 - offset/scale","    { static unsigned long long* tr = nullptr; if (!tr && getenv(\"ESLAM_LEAN_TRACE\")) { cudaMalloc((void**)&tr, 148 * 32 * 16 * 8); } if (tr) cudaMemsetAsync(tr, 0, 148 * 32 * 16 * 8, ctx->stream); F->trace = tr; }\n    fc.c0x = 402653184.0 - P.map.ox * P.map.inv_sx;                       // 1.5 * 2^28 - offset/scale",1)
s=s.replace("""    CK(cudaEventRecord(ctx->events[1], ctx->stream));
    CK(cudaEventSynchronize(ctx->events[1]));
    timing_report(ctx);""","""    CK(cudaEventRecord(ctx->events[1], ctx->stream));
    CK(cudaEventSynchronize(ctx->events[1]));
    if (ctx->fast->trace) { std::vector<unsigned long long> h(148 * 32 * 16); cudaMemcpy(h.data(), ctx->fast->trace, h.size() * 8, cudaMemcpyDeviceToHost); FILE* f = fopen("gpurun_out/lean_trace.bin", "wb"); if (f) { fwrite(h.data(), 8, h.size(), f); fclose(f); } }
    timing_report(ctx);""",1)
open(p,'w').write(s)
