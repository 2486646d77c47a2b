#!/bin/bash
# Records what the GPU box offers for running the C# reference itself ("Next round" item 1): any .NET / mono
# runtime or compiler, other managed toolchains, and the host's cores.  Output: gpurun_out/r02_probe/probe.txt .
#   gpurun --timeout 600 -- 'bash tools/probe_box.sh'
O=gpurun_out/r02_probe
mkdir -p $O
{
	echo "== date"; date -u
	echo "== uname"; uname -a
	echo "== cpu"; nproc; lscpu | head -25
	echo "== mem"; free -g | head -3
	echo "== gpu"; nvidia-smi -L; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv
	echo "== managed toolchains on PATH"
	for t in dotnet mono mcs csc msbuild xbuild pwsh nuget java javac go node rustc cargo; do
		p=$(command -v $t 2>/dev/null); echo "$t: ${p:-not found}"
	done
	echo "== dotnet --info"; dotnet --info 2>&1 | head -20
	echo "== mono --version"; mono --version 2>&1 | head -3
	echo "== well-known install dirs"
	ls -d /usr/share/dotnet /usr/lib/dotnet /usr/local/share/dotnet /opt/dotnet /root/.dotnet /usr/lib/mono /opt/mono 2>&1
	echo "== runtime files anywhere on the box (libcoreclr, libmono, System.Private.CoreLib, *.nupkg)"
	find / -xdev \( -name 'libcoreclr*' -o -name 'libmono*' -o -name 'System.Private.CoreLib*' -o -name 'mscorlib.dll' -o -name '*.nupkg' -o -name 'libhostfxr*' \) 2>/dev/null | head -20
	echo "(end of list)"
	echo "== python packages that could host the CLR"; python -c "import clr" 2>&1 | tail -1; python -c "import pythonnet" 2>&1 | tail -1
	echo "== /root/reference present?"; ls -d /root/reference 2>&1
} > $O/probe.txt 2>&1
cat $O/probe.txt
