# run_reference.jl -- times the UNMODIFIED GridapMHD.jl on the benchmark configuration of bench.py, for anyone with a
# Julia toolchain (none exists in the build image, so this script is shipped un-executed; SURVEY.md section 8(d)).
#
#   julia --project=/path/to/GridapMHD.jl baseline/run_reference.jl                      # sequential
#   mpiexec -np 4 julia --project=/path/to/GridapMHD.jl baseline/run_reference.jl 2 2    # MPI, np = (2,2,1)
#
# What is timed: `time_residual` and `time_jacobian`, the PTimer sections of GridapMHD.main (src/main.jl:157-164) that
# wrap `residual(op,xh)` / `jacobian(op,xh)` -- the same work bench.py's step (residual_and_jacobian!) does -- on
# Hunt nc=(64,64), Ha=1000 (hunt.jl:40-83 keyword names), assembled at the random initial state of main.jl:137-141.
# Output: one JSON line in the format of `bench.py --impl reference`.
using GridapMHD
using BSON
using Printf

px = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 0
py = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 0
path = mktempdir()
kw = (; nc=(64, 64), B=(0.0, 1000.0, 0.0), solve=false, res_assemble=true, jac_assemble=true, vtk=false,
      title="bench", path=path, verbose=false)
if px > 0
  GridapMHD.hunt(; backend=:mpi, np=(px, py, 1), kw...)
else
  GridapMHD.hunt(; kw...)
end
f = joinpath(path, "bench_r1.bson")
if isfile(f)   # written by the main rank only (hunt.jl:30-35)
  info = BSON.load(f)
  ncells = info[:ncells]
  t = info[:time_residual] + info[:time_jacobian]
  @printf("{\"impl\": \"reference\", \"metric\": \"mhd_assembly_jacobian_plus_residual\", \"value\": %.6g, \"unit\": \"Mcells/s\", ",
          ncells / t / 1e6)
  @printf("\"ncells\": %d, \"ndofs\": %d, \"time_residual_s\": %.6g, \"time_jacobian_s\": %.6g, \"ranks\": %d, \"kind\": \"reference\"}\n",
          ncells, info[:ndofs], info[:time_residual], info[:time_jacobian], max(px * py, 1))
end
