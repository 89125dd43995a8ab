# FermiB200.jl -- Julia glue that puts libfermi_pt_b200.so behind Fermi.jl's RCCSD(T) entry points.
#
# NOT EXECUTED IN THE BUILD IMAGE (no Julia there); it is the binding a Fermi.jl maintainer adds.  The same C ABI
# (include/fermi_pt_b200.h) is exercised from Python/ctypes by tests/ and bench.py.
#
# Usage (inside a Fermi.jl session):
#     include("FermiB200.jl")                      # defines FermiB200.B200 <: Fermi.CoupledCluster.RpTAlgorithm, registers pt_alg = 4
#     @set pt_alg 4
#     @energy ccsd(t)                              # -> RCCSDpT() -> RCCSDpT(FermiB200.B200()) -> RCCSDpT(ccsd, moints, FermiB200.B200())
#  or, warm:  @energy cc, moints => ccsd(t)   /   Fermi.CoupledCluster.RCCSDpT(cc, moints, FermiB200.B200())
#  FERMI_PT_B200_NGPU=8 makes the one handle drive GPUs 0..7 (sharded upload, NCCL all-gather, one scalar all-reduce).
#
# Mirrors src/Methods/CoupledCluster/PerturbativeTriples/ijk.jl:4-20 (the three overloads) and replaces
# ijk.jl:20-150 by one ccall.

module FermiB200

using Fermi
using Fermi.Options
using Fermi.Integrals: IntegralHelper, AbstractERI, AbstractDFERI, Chonky
using Fermi.Orbitals: AtomicOrbitals
using Fermi.Orbitals: AbstractRestrictedOrbitals
import Fermi.CoupledCluster: RCCSD, RCCSDpT, RpTAlgorithm, get_rpt_alg, ijk, ijk2, abc
import Fermi: output, Molecule, FermiException

const LIB = get(ENV, "FERMI_PT_B200_LIB", joinpath(@__DIR__, "..", "libfermi_pt_b200.so"))

struct B200 <: RpTAlgorithm end

# mirror of fpt_stats (include/fermi_pt_b200.h)
struct FptStats
    upload_ms::Cdouble; kernel_ms::Cdouble; total_ms::Cdouble; flops::Cdouble; h2d_bytes::Cdouble
    n_items::Clonglong; n_triplets::Clonglong; n_launches::Cint; n_sm::Cint
end

const HANDLE = Ref{Ptr{Cvoid}}(C_NULL)

function handle()
    if HANDLE[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        # FERMI_PT_B200_NGPU=8 -> one handle driving GPUs 0..7 (NCCL broadcast of the operands, one scalar all-reduce)
        ngpu = parse(Int, get(ENV, "FERMI_PT_B200_NGPU", "1"))
        devs = Cint.(collect(0:ngpu-1))
        rc = ccall((:fpt_create, LIB), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), ngpu, devs, h)
        rc == 0 || throw(FermiException(unsafe_string(ccall((:fpt_last_error, LIB), Cstring, ()))))
        HANDLE[] = h[]
        atexit(() -> ccall((:fpt_destroy, LIB), Cint, (Ptr{Cvoid},), HANDLE[]))
    end
    return HANDLE[]
end

check(rc) = rc == 0 || throw(FermiException(unsafe_string(ccall((:fpt_last_error, LIB), Cstring, ()))))

# pt_alg = 4 selects the B200 path (PerturbativeTriples.jl:3-11 holds the list as a local literal, so extend it here)
function Fermi.CoupledCluster.get_rpt_alg()
    implemented = [ijk(), ijk2(), abc(), B200()]
    N = Options.get("pt_alg")
    try
        return implemented[N]
    catch BoundsError
        throw(FermiException("implementation number $N not available for RCCSD(T)."))
    end
end

function RCCSDpT(Alg::B200)                                   # ijk.jl:4-10
    val = Options.get("return_ints")
    Options.set("return_ints", true)
    ccsd, moints = RCCSD()
    Options.set("return_ints", val)
    return RCCSDpT(ccsd, moints, Alg)
end

function RCCSDpT(mol::Molecule, Alg::B200)                    # ijk.jl:12-18
    val = Options.get("return_ints")
    Options.set("return_ints", true)
    ccsd, moints = RCCSD(mol)
    Options.set("return_ints", val)
    return RCCSDpT(ccsd, moints, Alg)
end

dense(A) = A isa Array{Float64} ? A : collect(Float64, A)
dense32(A) = A isa Array{Float32} ? A : collect(Float32, A)

function RCCSDpT(ccsd::RCCSD, moints::IntegralHelper{T,E,O}, Alg::B200) where {T<:AbstractFloat,
                                                                              E<:AbstractERI,O<:AbstractRestrictedOrbitals}
    # `@set precision single` (T = Float32, IntegralHelper.jl:58-68): cached blocks / B factors go down in 4-byte form through the f32
    # entry points (widened on the GPU); the correction is evaluated in FP64 -- the exact (T) energy of the rounded inputs, which the
    # reference's Float32 loops only approximate; the result struct keeps the caller's T
    output("\n   • Perturbative Triples Started\n")
    output("   - Contraction Engine: B200 DMMA (libfermi_pt_b200)")
    if T === Float32 && (haskey(moints.cache, "OVVV") || moints.eri_type isa AbstractDFERI)
        o, v = size(ccsd.T1)
        Et = Ref{Cdouble}(0.0)
        st = Ref{FptStats}()
        output("Computing energy contribution from occupied orbitals:")
        t = @elapsed begin
            T1 = dense32(ccsd.T1); T2 = dense32(ccsd.T2); fo = dense32(moints["Fii"]); fv = dense32(moints["Faa"])
            if haskey(moints.cache, "OVVV")
                check(ccall((:fpt_triples_conv_f32, LIB), Cint,
                            (Ptr{Cvoid}, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
                             Ref{Cdouble}, Ref{FptStats}),
                            handle(), o, v, T1, T2, dense32(moints["OVVV"]), dense32(moints["OOOV"]), dense32(moints["OVOV"]), fo, fv, Et, st))
            else
                BOV = dense32(moints["BOV"])
                check(ccall((:fpt_triples_df_f32, LIB), Cint,
                            (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat},
                             Ref{Cdouble}, Ref{FptStats}),
                            handle(), o, v, size(BOV, 1), T1, T2, dense32(moints["BOO"]), BOV, dense32(moints["BVV"]), fo, fv, Et, st))
            end
        end
        output("Finished in {:5.5f} s", t)
        output("Final (T) contribution: {:15.10f}", Et[])
        output("CCSD(T) energy:         {:15.10f}", Et[] + ccsd.energy)
        return RCCSDpT{T}(ccsd, T(Et[] + ccsd.energy), T(Et[]))
    end
    T1 = dense(ccsd.T1); T2 = dense(ccsd.T2)
    o, v = size(T1)
    fo = dense(moints["Fii"]); fv = dense(moints["Faa"])
    Et = Ref{Cdouble}(0.0)
    st = Ref{FptStats}()
    output("Computing energy contribution from occupied orbitals:")
    t = @elapsed begin
        if moints.eri_type isa AbstractDFERI && !haskey(moints.cache, "OVVV")
            # DF: hand over B factors ([Q,.,.], Q fastest, DFERI.jl:15-69); (ia|bd) etc. are assembled on the GPU
            BOO = dense(moints["BOO"]); BOV = dense(moints["BOV"]); BVV = dense(moints["BVV"])
            naux = size(BOV, 1)
            check(ccall((:fpt_triples_df, LIB), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                         Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{FptStats}),
                        handle(), o, v, naux, T1, T2, BOO, BOV, BVV, fo, fv, Et, st))
        elseif moints.eri_type isa Chonky && !haskey(moints.cache, "OVVV") && T === Float64 && get(ENV, "FERMI_PT_B200_AO", "dense") == "sparse"
            # opt-in (FERMI_PT_B200_AO=sparse; Float64 only -- the reference's sparse builder does not support single precision):
            # hand the GPU the *sparse* AO list (what IntegralHelper.jl:88-90 gives an AO helper by default) instead of letting
            # Sparse.jl:78-151,236-393 scatter and contract it on the CPU.  Not the default because compute!(::IntegralHelper{T,Chonky})
            # itself builds its AO helper with eri_type = I.eri_type, i.e. dense (ROIntegrals.jl:1-7): the branch below mirrors that.
            basis = moints.orbitals.basis                          # ROIntegrals.jl:3
            aoorbs = AtomicOrbitals(moints.molecule, basis)
            aoints = IntegralHelper{Float64}(molecule=moints.molecule, orbitals=aoorbs, basis=basis, eri_type=Fermi.Integrals.SparseERI())
            eri = aoints["ERI"]                                   # FermiSparse{Float64,Int16,4}: .indexes (zero-based), .data
            C = moints.orbitals.C
            core = Options.get("drop_occ"); inac = Options.get("drop_vir")
            ndocc = moints.molecule.Nα; nbf = size(C, 1)
            Co = dense(C[:, (1+core):ndocc]); Cv = dense(C[:, (ndocc+1):(nbf-inac)])
            Ti = eltype(eltype(eri.indexes))
            check(ccall((:fpt_triples_ao_sparse, LIB), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Clonglong, Ptr{Cvoid}, Cint, Ptr{Cdouble},
                         Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{FptStats}),
                        handle(), nbf, o, v, T1, T2, length(eri.data), eri.indexes, sizeof(Ti), eri.data, Co, Cv, fo, fv, Et, st))
        elseif moints.eri_type isa Chonky && !haskey(moints.cache, "OVVV")
            # dense AO integrals and no cached (ov|vv): instead of compute_OVVV!/OOOV!/OVOV! on the CPU (Chonky.jl:28-114) hand
            # the AO tensor and the orbital blocks over; the AO helper is built exactly as ROIntegrals.jl:1-7 builds it
            basis = moints.orbitals.basis                          # ROIntegrals.jl:3
            aoorbs = AtomicOrbitals(moints.molecule, basis)
            aoints = IntegralHelper{T}(molecule=moints.molecule, orbitals=aoorbs, basis=basis, eri_type=moints.eri_type)
            AOERI = dense(aoints["ERI"])
            C = moints.orbitals.C
            core = Options.get("drop_occ"); inac = Options.get("drop_vir")
            ndocc = moints.molecule.Nα; nbf = size(C, 1)
            Co = dense(C[:, (1+core):ndocc]); Cv = dense(C[:, (ndocc+1):(nbf-inac)])       # Chonky.jl:38-41
            check(ccall((:fpt_triples_ao, LIB), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                         Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{FptStats}),
                        handle(), nbf, o, v, T1, T2, AOERI, Co, Cv, fo, fv, Et, st))
        else
            OVVV = dense(moints["OVVV"]); OOOV = dense(moints["OOOV"]); OVOV = dense(moints["OVOV"])
            check(ccall((:fpt_triples_conv, LIB), Cint,
                        (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                         Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{FptStats}),
                        handle(), o, v, T1, T2, OVVV, OOOV, OVOV, fo, fv, Et, st))
        end
    end
    output("Finished in {:5.5f} s", t)
    output("Final (T) contribution: {:15.10f}", Et[])
    output("CCSD(T) energy:         {:15.10f}", Et[] + ccsd.energy)
    return RCCSDpT{T}(ccsd, T(Et[] + ccsd.energy), T(Et[]))  # ijk.jl:149
end

# ---- beside the (T) path (SURVEY 8f): the DF-CCSD particle-particle ladder and the DF-MP2 energy on the GPU ------------------------
# ladder_df!(newT2, T1, T2, moints) does what cc_update_T2_v4_term!(newT2, T1, T2, moints::IntegralHelper{T,<:AbstractDFERI}, ::RCCSDa)
# does (RCCSDHelper.jl:204-220), mp2_df(ints) what RMP2_energy(ints::IntegralHelper{T,<:AbstractDFERI,RHFOrbitals}, alg) does
# (RMP2a.jl:91-143).  FermiB200.offload_ccsd_ladder!() / offload_mp2!() redefine those two reference methods to call them.
function ladder_df!(newT2::Array{Float64,4}, T1, T2, moints::IntegralHelper)
    o, v = size(T1)
    BVV = dense(moints["BVV"])
    st = Ref{FptStats}()
    check(ccall((:fpt_ccsd_ladder_df, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{FptStats}),
                handle(), o, v, size(BVV, 1), dense(T1), dense(T2), BVV, newT2, st))
    return newT2
end

function mp2_df(ints::IntegralHelper)
    BOV = dense(ints["BOV"]); fo = dense(ints["Fii"]); fv = dense(ints["Faa"])
    E = Ref{Cdouble}(0.0)
    st = Ref{FptStats}()
    check(ccall((:fpt_mp2_df, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Cdouble}, Ref{FptStats}),
                handle(), length(fo), length(fv), size(BOV, 1), BOV, fo, fv, E, st))
    return E[]
end

function offload_ccsd_ladder!()
    @eval function Fermi.CoupledCluster.cc_update_T2_v4_term!(newT2::Array{Float64,4}, T1::AbstractArray{Float64,2}, T2::AbstractArray{Float64,4},
                                                              moints::IntegralHelper{Float64,E,O}, alg::Fermi.CoupledCluster.RCCSDa) where {
                                                              E<:AbstractDFERI,O<:AbstractRestrictedOrbitals}
        FermiB200.ladder_df!(newT2, T1, T2, moints)
    end
end

function offload_mp2!()
    @eval function Fermi.MollerPlesset.RMP2_energy(ints::IntegralHelper{Float64,<:AbstractDFERI,Fermi.Orbitals.RHFOrbitals},
                                                   Alg::Fermi.MollerPlesset.RMP2Algorithm)
        output(" Computing DF-MP2 Energy!")
        return FermiB200.mp2_df(ints)
    end
end

end # module
