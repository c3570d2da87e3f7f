# Reference CPU baseline for BASELINE config 2, for anyone with a Julia runtime (none exists in
# the build image, so this script ships UNEXECUTED; bench.py reports the C++ restatement instead).
#   julia -t auto scripts/julia_reference_baseline.jl [batch]
using PeriodicSchurDecompositions, Random
function main(B)
    n, p = 32, 8
    Random.seed!(1234)
    As = [[rand(n, n) for _ in 1:p] for _ in 1:B]
    pschur!([copy(a) for a in As[1]], :R; wantZ = false, wantT = false)  # compile
    t = @elapsed Threads.@threads for b in 1:B
        pschur!(As[b], :R; wantZ = false, wantT = false)
    end
    println("threads=$(Threads.nthreads()) batch=$B  $(B / t) problems/s")
end
main(length(ARGS) > 0 ? parse(Int, ARGS[1]) : 2000)
