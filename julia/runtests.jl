# runtests.jl — runs the REFERENCE's own test bodies against the CUDAArray backend, the way test/debug_array/*.jl and
# test/mpi_array/*.jl do for the other two backends (test/debug_array/p_vector_tests.jl: `with_debug(p_vector_tests)`).
#
#   PA_REFERENCE=/path/to/PartitionedArrays.jl PA_B200_LIB=/path/to/libpa_b200.so PA_B200_NGPUS=4 \
#       julia --project=$PA_REFERENCE -t 8 julia/runtests.jl
#
# STATUS: unverified — the build image has no Julia toolchain.  Part 1 exercises the AbstractArray contract of the backend
# with the unmodified reference tests (index partitions, exchange, PVector / PSparseMatrix algebra on host items through the
# backend's `map`); part 2 repeats the hot-path checks of those tests on device-resident PVector / PSparseMatrix objects
# (to_device) and compares with the host results of part 1's code path.
using Test, LinearAlgebra, SparseArrays
using PartitionedArrays

const REF = get(ENV, "PA_REFERENCE", joinpath(@__DIR__, "..", "..", "reference"))
include(joinpath(@__DIR__, "cuda_array.jl"))
using .PartitionedArraysB200

for f in ("primitives_tests.jl", "p_range_tests.jl", "p_vector_tests.jl", "p_sparse_matrix_tests.jl", "gallery_tests.jl", "fdm_example.jl", "fem_example.jl")
    include(joinpath(REF, "test", f))
end

@testset "CUDAArray backend: reference test bodies (host items through the backend's map/exchange)" begin
    with_cuda(primitives_tests)
    with_cuda(p_range_tests)
    with_cuda(p_vector_tests)
    with_cuda(p_sparse_matrix_tests)
    with_cuda(gallery_tests)
    with_cuda(fdm_example)
    with_cuda(fem_example)
end

@testset "CUDAArray backend: device-resident hot path vs the host path" begin
    with_cuda() do distribute
        np = parse(Int, get(ENV, "PA_B200_NGPUS", "1"))
        ranks = distribute(LinearIndices((np,)))
        # test/p_sparse_matrix_tests.jl:207-248: A = 2I, x = 3 => A*x == 6 on own values, and on ghosts after consistent!
        n = 10 * np
        rows = uniform_partition(ranks, n)
        I, J, V = map(rows) do r
            g = collect(own_to_global(r)); g, copy(g), fill(2.0, length(g))
        end |> tuple_of_arrays
        A = psparse(sparsecsr, I, J, V, rows, rows; assembled = true) |> fetch
        x = pfill(3.0, partition(axes(A, 2)))
        y = A * x
        dA, dx = to_device(A), to_device(x)
        dy = dA * dx
        @test to_host(dy) == y
        # the gallery operator (src/gallery.jl:12-98): mul!, dot, norm, consistent!, assemble!, CG
        parts = np == 1 ? (1, 1, 1) : (np, 1, 1)
        A = laplacian_fdm((8 * parts[1], 6, 5), parts, ranks) |> fetch
        x = pones(partition(axes(A, 2))); b = A * x
        dA, dx, db = to_device(A), to_device(x), to_device(b)
        dc = similar(db); mul!(dc, dA, dx)
        @test partition(to_host(dc)) == partition(b) || map(own_values(to_host(dc)), own_values(b)) do u, v; u == v end |> all
        @test dot(db, db) ≈ dot(b, b) rtol = 1e-12
        @test norm(db) ≈ norm(b) rtol = 1e-12
        consistent!(dx) |> wait; consistent!(x) |> wait
        @test partition(to_host(dx)).items == partition(x).items
        assemble!(dx) |> wait; assemble!(x) |> wait
        @test partition(to_host(dx)).items == partition(x).items
        x0 = pzeros(partition(axes(A, 2))); dx0 = to_device(x0)
        _, hist, r0, r, iters = ref_cg!(dx0, dA, to_device(b); tolerance = 1e-9, maxiter = 500)
        @test r / r0 <= 1e-9
        @test norm(to_host(dx0) - pones(partition(axes(A, 2)))) < 1e-5     # test/fdm_example.jl:128
    end
end
