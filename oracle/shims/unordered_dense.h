/*
 * unordered_dense.h -- TEST INFRASTRUCTURE (oracle shim).
 * The reference includes <unordered_dense.h> (ankerl, un-vendored submodule, absent here).
 * The hot-path TUs only need set/map membership (marchingCubes.cpp:54-55), so the std
 * containers are a behaviour-preserving stand-in (iteration order is never observed there).
 */
#ifndef UP_ORACLE_UNORDERED_DENSE_SHIM_H
#define UP_ORACLE_UNORDERED_DENSE_SHIM_H
#include <functional>
#include <unordered_map>
#include <unordered_set>
namespace ankerl::unordered_dense
{
	template<class T, class Enable = void>
	struct hash : std::hash<T>
	{};
	template<class K, class V, class H = hash<K>, class E = std::equal_to<K>>
	using map = std::unordered_map<K, V, H, E>;
	template<class K, class H = hash<K>, class E = std::equal_to<K>>
	using set = std::unordered_set<K, H, E>;
	template<class K, class V, class H = hash<K>, class E = std::equal_to<K>>
	using segmented_map = std::unordered_map<K, V, H, E>;
	template<class K, class H = hash<K>, class E = std::equal_to<K>>
	using segmented_set = std::unordered_set<K, H, E>;
}
#endif
