"""Data side of the path.  `get_dataset` mirrors datasets/__init__.py:8-46: the KITTI / nuScenes / Waymo readers and, on top of
them, the device-side batch construction (`train_siamese`, `train_motion`) or the raw tracklets (`test`)."""


def get_dataset(config, type='train', **kwargs):
    if config.dataset == 'kitti':
        from .kitti import kittiDataset
        data = kittiDataset(path=config.path, split=kwargs.get('split', 'train'), category_name=config.category_name,
                            coordinate_mode=config.coordinate_mode, preloading=config.preloading,
                            preload_offset=config.preload_offset if type != 'test' else -1)
    elif config.dataset == 'nuscenes':
        from .nuscenes_data import NuScenesDataset
        split = kwargs.get('split', 'train_track')
        data = NuScenesDataset(path=config.path, split=split, category_name=config.category_name, version=config.version,
                               key_frame_only=True if type != 'test' else config.key_frame_only, preloading=config.preloading,
                               preload_offset=config.preload_offset if type != 'test' else -1,
                               min_points=1 if split in [config.val_split, config.test_split] else -1, scenes=kwargs.get('scenes'))
    elif config.dataset == 'waymo':
        from .waymo_data import WaymoDataset
        data = WaymoDataset(path=config.path, split=kwargs.get('split', 'train'), category_name=config.category_name,
                            preloading=config.preloading, preload_offset=config.preload_offset, tiny=config.tiny)
    else:
        raise NotImplementedError(f"dataset '{config.dataset}'")
    if type == 'train_siamese':
        from .device_sampler import DeviceSiameseSampler
        return DeviceSiameseSampler(data.tracklets(), config, kwargs.get('device', 'cuda'))
    if type.lower() == 'train_motion':
        from .device_sampler import DeviceMotionSampler
        return DeviceMotionSampler(data.tracklets(), config, kwargs.get('device', 'cuda'))
    return data.tracklets()
